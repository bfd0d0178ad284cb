"""ctypes mirror of include/mmdgpu.h.

Only type definitions and array marshalling live here; loading the CUDA library is in
`simple_mmd_renderer_b200.lib`.  The same descriptor structs are handed to the test oracles
(oracle/) so that the reference, the restatement and the device path all see identical arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

# ---------------------------------------------------------------- constants (mmdgpu.h enums)
OK, ERR_INVALID_ARG, ERR_BAD_INDEX, ERR_UNSUPPORTED, ERR_CUDA, ERR_OOM, ERR_PARSE = 0, -1, -2, -3, -4, -5, -6

SKIN_BDEF1, SKIN_BDEF2, SKIN_BDEF4, SKIN_SDEF, SKIN_QDEF = 0, 1, 2, 3, 4
BONE_HAS_IK, BONE_APPEND_ROTATE, BONE_APPEND_TRANSLATE, BONE_POST_PHYSICS = 0x20, 0x100, 0x200, 0x1000
(MORPH_GROUP, MORPH_VERTEX, MORPH_BONE, MORPH_UV, MORPH_EXT_UV1, MORPH_EXT_UV2, MORPH_EXT_UV3, MORPH_EXT_UV4,
 MORPH_MATERIAL) = range(9)
LAYOUT_SOA_POS_NRM, LAYOUT_INTERLEAVED_SOKOL32 = 0, 1
STREAM_POSITION, STREAM_NORMAL, STREAM_INTERLEAVED, STREAM_SKIN_MATRIX, STREAM_UV = 0, 1, 2, 3, 4

(PLAN_SKIN_TYPE, PLAN_BONE_ID, PLAN_WEIGHT, PLAN_ORDER_PRE, PLAN_ORDER_POST, PLAN_OP_KIND, PLAN_OP_BONE,
 PLAN_OP_WAVE, PLAN_WAVE_BEGIN, PLAN_WAVE_OPS, PLAN_IK_FIX_TYPE, PLAN_IK_EULER_ORDER, PLAN_APP_SLOT_MORPH,
 PLAN_APP_SLOT_PARENT, PLAN_APP_SLOT_MULT, PLAN_CSR_ROW_PTR, PLAN_CSR_SLOT, PLAN_CSR_OFFSET, PLAN_BEZIER_UNUSED,
 PLAN_WAVE_PHASE_SPLIT, PLAN_TILE_ORIG, PLAN_TILE_TYPE, PLAN_TILE_LOCAL_ID, PLAN_TILE_BONE_BEGIN, PLAN_TILE_BONES,
 PLAN_ELL_BASE, PLAN_ELL_ROUNDS, PLAN_ELL_SLOT, PLAN_ELL_OFFSET, PLAN_POSITION, PLAN_NORMAL, PLAN_UV,
 PLAN_BONE_STATIC, PLAN_IK_DESC, PLAN_IK_LINK, PLAN_BONE_MORPH, PLAN_MATERIAL_MORPH_ROW,
 PLAN_MATERIAL_MORPH, PLAN_IK_IMAGE, PLAN_IK_IMAGE_BONES, PLAN_IK_IMAGE_WRITTEN, PLAN_IK_IMAGE_STATIC,
 PLAN_IK_IMAGE_LINK_SLOTS, PLAN_IK_IMAGE_MORPH_SLOTS, PLAN_IK_IMAGE_DESC, PLAN_IK_IMAGE_LINKS) = range(46)

PLAN_DTYPES = {
    PLAN_SKIN_TYPE: np.uint8, PLAN_BONE_ID: np.uint16, PLAN_WEIGHT: np.float32, PLAN_ORDER_PRE: np.int32,
    PLAN_ORDER_POST: np.int32, PLAN_OP_KIND: np.uint8, PLAN_OP_BONE: np.int32, PLAN_OP_WAVE: np.int32,
    PLAN_WAVE_BEGIN: np.int32, PLAN_WAVE_OPS: np.int32, PLAN_IK_FIX_TYPE: np.uint8, PLAN_IK_EULER_ORDER: np.uint8,
    PLAN_APP_SLOT_MORPH: np.int32, PLAN_APP_SLOT_PARENT: np.int32, PLAN_APP_SLOT_MULT: np.float32,
    PLAN_CSR_ROW_PTR: np.uint32, PLAN_CSR_SLOT: np.uint32, PLAN_CSR_OFFSET: np.float32,
    PLAN_WAVE_PHASE_SPLIT: np.int32,
    PLAN_TILE_ORIG: np.uint16, PLAN_TILE_TYPE: np.uint8, PLAN_TILE_LOCAL_ID: np.uint16,
    PLAN_TILE_BONE_BEGIN: np.uint32, PLAN_TILE_BONES: np.uint16, PLAN_ELL_BASE: np.uint32,
    PLAN_ELL_ROUNDS: np.uint32, PLAN_ELL_SLOT: np.uint32, PLAN_ELL_OFFSET: np.float32,
    PLAN_POSITION: np.float32, PLAN_NORMAL: np.float32, PLAN_UV: np.float32, PLAN_BONE_STATIC: np.uint8,
    PLAN_IK_DESC: np.uint8, PLAN_IK_LINK: np.uint8, PLAN_BONE_MORPH: np.uint8,
    PLAN_MATERIAL_MORPH_ROW: np.int32, PLAN_MATERIAL_MORPH: np.uint8,
    PLAN_IK_IMAGE: np.uint8, PLAN_IK_IMAGE_BONES: np.int32, PLAN_IK_IMAGE_WRITTEN: np.uint8, PLAN_IK_IMAGE_STATIC: np.uint8,
    PLAN_IK_IMAGE_LINK_SLOTS: np.int32, PLAN_IK_IMAGE_MORPH_SLOTS: np.int32, PLAN_IK_IMAGE_DESC: np.uint8,
    PLAN_IK_IMAGE_LINKS: np.uint8,
}

(ANIM_BONE_KEY_BEGIN, ANIM_BONE_KEY_COUNT, ANIM_BONE_TRACKED, ANIM_KEY_FRAME, ANIM_KEY_T, ANIM_KEY_R, ANIM_KEY_CURVE,
 ANIM_TABLES, ANIM_MORPH_KEY_BEGIN, ANIM_MORPH_KEY_COUNT, ANIM_MORPH_TRACKED, ANIM_MKEY_FRAME,
 ANIM_MKEY_WEIGHT) = range(13)
ANIM_DTYPES = {
    ANIM_BONE_KEY_BEGIN: np.uint32, ANIM_BONE_KEY_COUNT: np.uint32, ANIM_BONE_TRACKED: np.uint8,
    ANIM_KEY_FRAME: np.uint32, ANIM_KEY_T: np.float32, ANIM_KEY_R: np.float32, ANIM_KEY_CURVE: np.uint32,
    ANIM_TABLES: np.float32, ANIM_MORPH_KEY_BEGIN: np.uint32, ANIM_MORPH_KEY_COUNT: np.uint32,
    ANIM_MORPH_TRACKED: np.uint8, ANIM_MKEY_FRAME: np.uint32, ANIM_MKEY_WEIGHT: np.float32,
}

# ---------------------------------------------------------------- numpy record dtypes (AoS pools)
VERTEX_MORPH_ENTRY = np.dtype([("vertex", "<u4"), ("offset", "<f4", (3,))])
UV_MORPH_ENTRY = np.dtype([("vertex", "<u4"), ("offset", "<f4", (4,))])
BONE_MORPH_ENTRY = np.dtype([("bone", "<u4"), ("translation", "<f4", (3,)), ("rotation", "<f4", (4,))])
GROUP_MORPH_ENTRY = np.dtype([("morph", "<u4"), ("rate", "<f4")])
MATERIAL_FIELDS = 28
MATERIAL_MUL, MATERIAL_ADD = 0, 1
MATERIAL_MORPH_ENTRY = np.dtype([("material", "<i4"), ("method", "<u4"), ("value", "<f4", (MATERIAL_FIELDS,))])
assert MATERIAL_MORPH_ENTRY.itemsize == 120
BONE_KEY = np.dtype([("frame", "<u4"), ("translation", "<f4", (3,)), ("rotation", "<f4", (4,)),
                     ("interp", "i1", (4, 4))])
MORPH_KEY = np.dtype([("frame", "<u4"), ("weight", "<f4")])
assert VERTEX_MORPH_ENTRY.itemsize == 16 and UV_MORPH_ENTRY.itemsize == 20 and BONE_MORPH_ENTRY.itemsize == 32
assert GROUP_MORPH_ENTRY.itemsize == 8 and BONE_KEY.itemsize == 48 and MORPH_KEY.itemsize == 8

_P = C.c_void_p


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_uint32),
        ("position", _P), ("normal", _P), ("uv", _P), ("skin_type", _P), ("bone_id", _P), ("weight", _P),
        ("sdef_c", _P), ("sdef_r0", _P), ("sdef_r1", _P),
        ("n_bones", C.c_uint32),
        ("bone_position", _P), ("bone_parent", _P), ("bone_transform_level", _P), ("bone_flags", _P),
        ("bone_append_parent", _P), ("bone_append_ratio", _P),
        ("ik_target", _P), ("ik_iterations", _P), ("ik_angle_limit", _P), ("ik_link_begin", _P),
        ("ik_link_count", _P),
        ("n_ik_links", C.c_uint32),
        ("ik_link_bone", _P), ("ik_link_has_limit", _P), ("ik_link_lo", _P), ("ik_link_hi", _P),
        ("n_morphs", C.c_uint32),
        ("morph_type", _P), ("morph_entry_begin", _P), ("morph_entry_count", _P),
        ("n_vertex_morph_entries", C.c_uint32), ("vertex_morph_entries", _P),
        ("n_uv_morph_entries", C.c_uint32), ("uv_morph_entries", _P),
        ("n_bone_morph_entries", C.c_uint32), ("bone_morph_entries", _P),
        ("n_group_morph_entries", C.c_uint32), ("group_morph_entries", _P),
        ("n_materials", C.c_uint32),
        ("n_material_morph_entries", C.c_uint32), ("material_morph_entries", _P),
    ]


class AnimDesc(C.Structure):
    _fields_ = [
        ("n_bone_tracks", C.c_uint32),
        ("bone_track_bone", _P), ("bone_track_key_begin", _P), ("bone_track_key_count", _P),
        ("n_bone_keys", C.c_uint32), ("bone_keys", _P),
        ("n_morph_tracks", C.c_uint32),
        ("morph_track_morph", _P), ("morph_track_key_begin", _P), ("morph_track_key_count", _P),
        ("n_morph_keys", C.c_uint32), ("morph_keys", _P),
    ]


class Options(C.Structure):
    _fields_ = [("extensions", C.c_uint32), ("reserved", C.c_uint32 * 7)]


# field -> (dtype, per-item width relative to the count field)
_MODEL_ARRAYS = {
    "position": (np.float32, "n_vertices", 3), "normal": (np.float32, "n_vertices", 3),
    "uv": (np.float32, "n_vertices", 2), "skin_type": (np.uint8, "n_vertices", 1),
    "bone_id": (np.int32, "n_vertices", 4), "weight": (np.float32, "n_vertices", 4),
    "sdef_c": (np.float32, "n_vertices", 3), "sdef_r0": (np.float32, "n_vertices", 3),
    "sdef_r1": (np.float32, "n_vertices", 3),
    "bone_position": (np.float32, "n_bones", 3), "bone_parent": (np.int32, "n_bones", 1),
    "bone_transform_level": (np.int32, "n_bones", 1), "bone_flags": (np.uint16, "n_bones", 1),
    "bone_append_parent": (np.int32, "n_bones", 1), "bone_append_ratio": (np.float32, "n_bones", 1),
    "ik_target": (np.int32, "n_bones", 1), "ik_iterations": (np.int32, "n_bones", 1),
    "ik_angle_limit": (np.float32, "n_bones", 1), "ik_link_begin": (np.uint32, "n_bones", 1),
    "ik_link_count": (np.uint32, "n_bones", 1),
    "ik_link_bone": (np.int32, "n_ik_links", 1), "ik_link_has_limit": (np.uint8, "n_ik_links", 1),
    "ik_link_lo": (np.float32, "n_ik_links", 3), "ik_link_hi": (np.float32, "n_ik_links", 3),
    "morph_type": (np.uint8, "n_morphs", 1), "morph_entry_begin": (np.uint32, "n_morphs", 1),
    "morph_entry_count": (np.uint32, "n_morphs", 1),
    "vertex_morph_entries": (VERTEX_MORPH_ENTRY, "n_vertex_morph_entries", 1),
    "uv_morph_entries": (UV_MORPH_ENTRY, "n_uv_morph_entries", 1),
    "bone_morph_entries": (BONE_MORPH_ENTRY, "n_bone_morph_entries", 1),
    "group_morph_entries": (GROUP_MORPH_ENTRY, "n_group_morph_entries", 1),
    "material_morph_entries": (MATERIAL_MORPH_ENTRY, "n_material_morph_entries", 1),
}
_MODEL_COUNTS = ["n_vertices", "n_bones", "n_ik_links", "n_morphs", "n_vertex_morph_entries",
                 "n_uv_morph_entries", "n_bone_morph_entries", "n_group_morph_entries", "n_materials",
                 "n_material_morph_entries"]

_ANIM_ARRAYS = {
    "bone_track_bone": (np.int32, "n_bone_tracks", 1), "bone_track_key_begin": (np.uint32, "n_bone_tracks", 1),
    "bone_track_key_count": (np.uint32, "n_bone_tracks", 1), "bone_keys": (BONE_KEY, "n_bone_keys", 1),
    "morph_track_morph": (np.int32, "n_morph_tracks", 1), "morph_track_key_begin": (np.uint32, "n_morph_tracks", 1),
    "morph_track_key_count": (np.uint32, "n_morph_tracks", 1), "morph_keys": (MORPH_KEY, "n_morph_keys", 1),
}
_ANIM_COUNTS = ["n_bone_tracks", "n_bone_keys", "n_morph_tracks", "n_morph_keys"]


def _fill(struct, arrays, counts, table, src):
    """Populate a ctypes struct from a dict of numpy arrays; returns the list that keeps them alive."""
    keep = []
    for c in counts:
        setattr(struct, c, int(src.get(c, 0)))
    for name, (dtype, count_field, width) in table.items():
        a = src.get(name)
        if a is None:
            setattr(struct, name, None)
            continue
        a = np.ascontiguousarray(a, dtype=dtype)
        need = int(getattr(struct, count_field)) * width
        if a.size != need:
            raise ValueError(f"{name}: expected {need} items, got {a.size}")
        keep.append(a)
        setattr(struct, name, a.ctypes.data if a.size else None)
    return keep


def model_desc(src: dict):
    """dict of numpy arrays (see synth.make_model) -> (ModelDesc, keepalive list)."""
    d = ModelDesc()
    keep = _fill(d, None, _MODEL_COUNTS, _MODEL_ARRAYS, src)
    return d, keep


def anim_desc(src: dict):
    d = AnimDesc()
    keep = _fill(d, None, _ANIM_COUNTS, _ANIM_ARRAYS, src)
    return d, keep
