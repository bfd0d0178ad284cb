"""The PMX / VMD byte-stream parsers pinned against libmmd's OWN readers (SURVEY 8f-2, 8b adapter).

oracle/_ref/reader_check loads a PMX + VMD pair with libmmd's PmxReader / VmdReader (L/reader/pmx_reader_impl.inl,
vmd_reader_impl.inl), flattens the resulting mmd::Model / mmd::Motion through include/mmdgpu_libmmd_adapter.hpp and
dumps the descriptors.  The host plan and the flattened motion built from that dump must be byte-identical to the
ones mmdgpu builds from the same files' bytes.  Runs only where the tool was built (the container that mounts the
reference); test infrastructure."""
import os
import struct
import subprocess

import numpy as np
import pytest

import pmxio
from conftest import synth_case
from simple_mmd_renderer_b200 import capi, synth
from simple_mmd_renderer_b200.poser import HostPlan

TOOL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "reader_check")
pytestmark = pytest.mark.skipif(not os.path.exists(TOOL), reason="oracle/_ref/reader_check not built (needs the reference headers)")


def _load_dump(path):
    raw = {}
    with open(path, "rb") as f:
        data = f.read()
    at = 0
    while at < len(data):
        (nl,) = struct.unpack_from("<I", data, at); at += 4
        name = data[at:at + nl].decode(); at += nl
        (nb,) = struct.unpack_from("<Q", data, at); at += 8
        raw[name] = data[at:at + nb]; at += nb
    scalars = {k: int(np.frombuffer(raw[k], np.uint32)[0]) for k in ("n_vertices", "n_bones", "n_ik_links", "n_morphs", "n_materials", "has_sdef")}
    model = {k: scalars[k] for k in ("n_vertices", "n_bones", "n_ik_links", "n_morphs", "n_materials")}
    for name, (dtype, _, _) in capi._MODEL_ARRAYS.items():
        if name.startswith("sdef_") and not scalars["has_sdef"]:
            continue
        model[name] = np.frombuffer(raw[name], dtype).copy()
    for pool in ("vertex", "uv", "bone", "group", "material"):
        model[f"n_{pool}_morph_entries"] = model[f"{pool}_morph_entries"].size
    motion = {}
    for name, (dtype, _, _) in capi._ANIM_ARRAYS.items():
        motion[name] = np.frombuffer(raw[name], dtype).copy()
    motion.update(n_bone_tracks=motion["bone_track_bone"].size, n_bone_keys=motion["bone_keys"].size,
                  n_morph_tracks=motion["morph_track_morph"].size, n_morph_keys=motion["morph_keys"].size)
    return model, motion


def _same(a: dict, b: dict, what: str):
    assert a.keys() == b.keys()
    for k in a:
        x, y = a[k], b[k]
        assert x.shape == y.shape, f"{what} array {k}: {x.shape} vs {y.shape}"
        np.testing.assert_array_equal(x.view(np.uint8) if x.dtype.kind == "f" else x,
                                      y.view(np.uint8) if y.dtype.kind == "f" else y, err_msg=f"{what} array {k}")


@pytest.mark.parametrize("name,materials,ascii_names", [("tiny", False, True), ("tiny_full", True, False), ("small", False, False),
                                                        ("ik_zoo", False, False)])
def test_byte_stream_parsers_match_libmmd_readers(tmp_path, name, materials, ascii_names):
    cfg, model, motion = synth_case(name)
    model = dict(model)
    ap = model["bone_append_parent"].copy()
    ap[(ap < 0) | (ap >= model["n_bones"])] = -1            # not representable in a narrow PMX index field
    model["bone_append_parent"] = ap
    st = model["skin_type"].copy()
    st[st == capi.SKIN_QDEF] = capi.SKIN_BDEF4              # libmmd's reader throws on QDEF (pmx_reader_impl.inl:95-98)
    model["skin_type"] = st
    if materials:
        model = synth.add_material_morphs(model)
        motion = synth.make_motion(cfg, model)
    pmxio.ASCII_NAMES = ascii_names
    try:
        pmx, vmd = pmxio.write_pmx(model), pmxio.write_vmd(motion)
    finally:
        pmxio.ASCII_NAMES = False
    (tmp_path / "m.pmx").write_bytes(pmx)
    (tmp_path / "m.vmd").write_bytes(vmd)
    r = subprocess.run([TOOL, str(tmp_path / "m.pmx"), str(tmp_path / "m.vmd"), str(tmp_path / "dump.bin")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    ref_model, ref_motion = _load_dump(tmp_path / "dump.bin")
    assert ref_model["n_vertices"] == model["n_vertices"] and ref_model["n_bones"] == model["n_bones"]
    assert ref_motion["n_bone_keys"] > 0 and (ref_motion["n_morph_keys"] > 0 or model["n_morphs"] == 0), \
        "libmmd joined no tracks: " + r.stdout

    for ext in (False, True):
        ours = HostPlan(pmx_bytes=pmx, extensions=ext)
        theirs = HostPlan(arrays=ref_model, extensions=ext)
        _same(theirs.arrays(), ours.arrays(), f"plan (extensions={ext})")
    want = theirs.anim_from_arrays(ref_motion, model["n_bones"], model["n_morphs"])
    got = ours.anim_from_vmd(vmd)
    _same(want, got, "flattened motion")
