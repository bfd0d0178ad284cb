"""Test-side writers: flat model / motion arrays -> PMX 2.0 / 2.1 and VMD byte streams (the layouts the reference's
readers consume: L/reader/pmx_reader_impl.inl:16-449, L/reader/vmd_reader_impl.inl:9-79).  Names mix ASCII, kanji,
hiragana and half-width katakana so that the CP932 <-> UTF-16 / UTF-8 name join is exercised."""
import struct

import numpy as np

from simple_mmd_renderer_b200 import capi

_KANA = ["センター", "上半身", "下半身", "左足", "右足", "左ひざ", "右ひざ", "頭", "首", "ｶﾀ", "左腕", "右腕"]


ASCII_NAMES = False     # libmmd's own readers only join ASCII names reliably under glibc (dwarf_impl.inl:205-232)


def bone_name(i: int) -> str:
    return f"bone_{i}" if ASCII_NAMES else f"{_KANA[i % len(_KANA)]}{i}"


def morph_name(i: int) -> str:
    return f"morph_{i}" if ASCII_NAMES else ["まばたき", "あ", "笑い", "ｳｨﾝｸ"][i % 4] + str(i)


def _idx_size(n: int) -> int:
    return 1 if n < 127 else 2 if n < 32767 else 4


def _pack_idx(v: int, size: int, signed: bool = True) -> bytes:
    if size == 1:
        return struct.pack("<B", v & 0xFF)
    if size == 2:
        return struct.pack("<H", v & 0xFFFF)
    return struct.pack("<i", v)


def _text(s: str, utf8: bool) -> bytes:
    b = s.encode("utf-8" if utf8 else "utf-16-le")
    return struct.pack("<i", len(b)) + b


def write_pmx(model: dict, utf8: bool = False, version: float = 2.0, extra_uv: int = 0) -> bytes:
    nv, nb, nm = int(model["n_vertices"]), int(model["n_bones"]), int(model["n_morphs"])
    vsz = 1 if nv < 256 else 2 if nv < 65536 else 4
    bsz, msz = _idx_size(nb), _idx_size(nm)
    tsz = matsz = rsz = 1
    out = [b"PMX ", struct.pack("<f", version), bytes([8, 1 if utf8 else 0, extra_uv, vsz, tsz, matsz, bsz, msz, rsz])]
    for s in ("model", "model_en", "comment", "comment_en"):
        out.append(_text(s, utf8))
    out.append(struct.pack("<i", nv))
    pos, nrm, uv = model["position"], model["normal"], model["uv"]
    st, bid, w = model["skin_type"], model["bone_id"], model["weight"]
    has_sdef = "sdef_c" in model and model["sdef_c"] is not None
    for i in range(nv):
        out.append(struct.pack("<8f", *pos[i], *nrm[i], *uv[i]))
        out.append(b"\0" * (16 * extra_uv))
        t = int(st[i])
        out.append(bytes([t]))
        if t == capi.SKIN_BDEF1:
            out.append(_pack_idx(int(bid[i, 0]), bsz))
        elif t == capi.SKIN_BDEF2:
            out.append(_pack_idx(int(bid[i, 0]), bsz) + _pack_idx(int(bid[i, 1]), bsz) + struct.pack("<f", w[i, 0]))
        elif t in (capi.SKIN_BDEF4, capi.SKIN_QDEF):
            out.append(b"".join(_pack_idx(int(bid[i, k]), bsz) for k in range(4)) + struct.pack("<4f", *w[i]))
        else:
            c = model["sdef_c"][i] if has_sdef else (0, 0, 0)
            r0 = model["sdef_r0"][i] if has_sdef else (0, 0, 0)
            r1 = model["sdef_r1"][i] if has_sdef else (0, 0, 0)
            out.append(_pack_idx(int(bid[i, 0]), bsz) + _pack_idx(int(bid[i, 1]), bsz) + struct.pack("<f", w[i, 0])
                       + struct.pack("<9f", *c, *r0, *r1))
        out.append(struct.pack("<f", 1.0))
    # faces: one triangle; textures: one; materials: one
    out.append(struct.pack("<i", 3) + b"".join(_pack_idx(0, vsz) for _ in range(3)))
    out.append(struct.pack("<i", 1) + _text("tex.png", utf8))
    n_mat = max(1, int(model.get("n_materials", 0)))
    out.append(struct.pack("<i", n_mat))
    for i in range(n_mat):
        out.append(_text(f"mat{i}", utf8) + _text(f"mat_en{i}", utf8) + b"\0" * 65
                   + _pack_idx(0, tsz) + _pack_idx(0xFF, tsz) + bytes([0, 1, 3]) + _text("memo", utf8)
                   + struct.pack("<i", 3 if i == 0 else 0))
    # bones
    out.append(struct.pack("<i", nb))
    flags = model["bone_flags"]
    for b in range(nb):
        out.append(_text(bone_name(b), utf8) + _text(f"bone{b}", utf8))
        out.append(struct.pack("<3f", *model["bone_position"][b]))
        p = int(model["bone_parent"][b])
        out.append(_pack_idx(p if p >= 0 else -1, bsz))
        out.append(struct.pack("<i", int(model["bone_transform_level"][b])))
        fl = int(flags[b]) | 0x0002 | (0x0001 if b % 2 else 0) | (0x0400 if b % 5 == 1 else 0) | (0x0800 if b % 7 == 2 else 0) \
            | (0x2000 if b % 11 == 3 else 0)
        out.append(struct.pack("<H", fl))
        out.append(_pack_idx(-1, bsz) if fl & 1 else struct.pack("<3f", 0, 1, 0))
        if fl & (capi.BONE_APPEND_ROTATE | capi.BONE_APPEND_TRANSLATE):
            ap = int(model["bone_append_parent"][b])
            # an append index beyond the bone count cannot be written with a narrow index: use -1 ("none")
            out.append(_pack_idx(ap if 0 <= ap < nb else -1, bsz) + struct.pack("<f", model["bone_append_ratio"][b]))
        if fl & 0x0400:
            out.append(struct.pack("<3f", 1, 0, 0))
        if fl & 0x0800:
            out.append(struct.pack("<6f", 1, 0, 0, 0, 0, 1))
        if fl & 0x2000:
            out.append(struct.pack("<i", 7))
        if fl & capi.BONE_HAS_IK:
            lb, lc = int(model["ik_link_begin"][b]), int(model["ik_link_count"][b])
            out.append(_pack_idx(int(model["ik_target"][b]), bsz) + struct.pack("<if", int(model["ik_iterations"][b]),
                                                                               model["ik_angle_limit"][b]))
            out.append(struct.pack("<i", lc))
            for l in range(lb, lb + lc):
                has = int(model["ik_link_has_limit"][l])
                out.append(_pack_idx(int(model["ik_link_bone"][l]), bsz) + bytes([has]))
                if has:
                    out.append(struct.pack("<6f", *model["ik_link_lo"][l], *model["ik_link_hi"][l]))
    # morphs
    out.append(struct.pack("<i", nm))
    mt, mb, mc = model["morph_type"], model["morph_entry_begin"], model["morph_entry_count"]
    for m in range(nm):
        t = int(mt[m])
        out.append(_text(morph_name(m), utf8) + _text(f"morph{m}", utf8) + bytes([1, t]))
        b0, n = int(mb[m]), int(mc[m])
        out.append(struct.pack("<i", n))
        if t == capi.MORPH_GROUP:
            for e in model["group_morph_entries"][b0:b0 + n]:
                out.append(_pack_idx(int(e["morph"]), msz) + struct.pack("<f", e["rate"]))
        elif t == capi.MORPH_VERTEX:
            for e in model["vertex_morph_entries"][b0:b0 + n]:
                out.append(_pack_idx(int(e["vertex"]), vsz) + struct.pack("<3f", *e["offset"]))
        elif t == capi.MORPH_BONE:
            for e in model["bone_morph_entries"][b0:b0 + n]:
                out.append(_pack_idx(int(e["bone"]), bsz) + struct.pack("<7f", *e["translation"], *e["rotation"]))
        elif capi.MORPH_UV <= t <= capi.MORPH_EXT_UV4:
            for e in model["uv_morph_entries"][b0:b0 + n]:
                out.append(_pack_idx(int(e["vertex"]), vsz) + struct.pack("<4f", *e["offset"]))
        elif t == capi.MORPH_MATERIAL:
            for e in model.get("material_morph_entries", ())[b0:b0 + n]:
                # "every material" is stored as -1, which ReadIndex returns as 255 for a 1-byte index
                out.append(_pack_idx(int(e["material"]), matsz) + bytes([int(e["method"])])
                           + struct.pack("<28f", *e["value"]))
    # display frames, rigid bodies, joints: empty
    out.append(struct.pack("<iii", 0, 0, 0))
    return b"".join(out)


def _sjis15(s: str) -> bytes:
    b = s.encode("cp932")
    assert len(b) <= 15, s
    return b + b"\0" * (15 - len(b))


def write_vmd(motion: dict, extra_unknown_tracks: bool = True) -> bytes:
    out = [b"Vocaloid Motion Data 0002".ljust(30, b"\0"), b"model".ljust(20, b"\0")]
    recs = []
    keys = motion["bone_keys"]
    for t in range(int(motion["n_bone_tracks"])):
        b = int(motion["bone_track_bone"][t])
        b0, n = int(motion["bone_track_key_begin"][t]), int(motion["bone_track_key_count"][t])
        for k in keys[b0:b0 + n]:
            blocks = b""
            for c in range(4):
                blk = bytearray(16)
                for q in range(4):
                    blk[4 * q] = int(k["interp"][c][q]) & 0xFF
                blocks += bytes(blk)
            recs.append(_sjis15(bone_name(b)) + struct.pack("<I3f4f", int(k["frame"]), *k["translation"], *k["rotation"]) + blocks)
    if extra_unknown_tracks:
        recs.append(_sjis15("存在しない") + struct.pack("<I3f4f", 3, 1, 2, 3, 0, 0, 0, 1) + b"\x14" * 64)
    out.append(struct.pack("<I", len(recs)) + b"".join(recs))
    mrecs = []
    mkeys = motion["morph_keys"]
    for t in range(int(motion["n_morph_tracks"])):
        m = int(motion["morph_track_morph"][t])
        b0, n = int(motion["morph_track_key_begin"][t]), int(motion["morph_track_key_count"][t])
        for k in mkeys[b0:b0 + n]:
            mrecs.append(_sjis15(morph_name(m)) + struct.pack("<If", int(k["frame"]), k["weight"]))
    out.append(struct.pack("<I", len(mrecs)) + b"".join(mrecs))
    out.append(struct.pack("<III", 0, 0, 0))   # camera, light, self-shadow sections
    return b"".join(out)
