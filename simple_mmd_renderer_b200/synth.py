"""Deterministic synthetic PMX/VMD-shaped inputs.

One generator feeds the libmmd reference harness, the C restatement and the CUDA library with the same
flat arrays (the dict layout is the one `capi.model_desc` / `capi.anim_desc` marshal).  Random numbers
come from numpy's PCG64 seeded with ``0x4D4D4400 + config id`` (+ instance for crowd clips); the draw
order is the statement order of this file.  Nothing here is on the product path: the library takes
whatever arrays a caller gives it.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

from . import capi

SEED_BASE = 0x4D4D4400
PI_F = np.float32(3.1415927)


@dataclass(frozen=True)
class SynthConfig:
    name: str
    config_id: int
    n_vertices: int
    n_bones: int
    n_vertex_morphs: int
    n_frames: int
    binding: str = "coherent"        # "coherent": bones from floor(i*NB/NV) +- 3 ; "wide": +- 24 (~50 bones per 512-vertex tile) ; "random": uniform
    sdef_qdef: bool = False          # tag 15 % of BDEF2 as SDEF, 10 % of BDEF4 as QDEF
    n_uv_morphs: int = 0
    n_group_morphs: int = 0
    n_bone_morphs: int = 0
    ik_chains: int = 0               # hip -> knee -> ankle(target) + IK bone parented to root
    post_physics_frac: float = 0.0
    stress: bool = False             # extra discontinuities for parity tests (see make_model / make_motion)
    morph_run_frac: float = 0.018
    morph_scatter_frac: float = 0.002


# BASELINE.json configs.  C4 = 512 instances of C1's model; C5 = C3's model over 10 k frames.
C1 = SynthConfig("C1", 1, 50_000, 200, 30, 300)
C2 = SynthConfig("C2", 2, 50_000, 200, 30, 300, sdef_qdef=True, n_uv_morphs=4, n_group_morphs=2,
                 n_bone_morphs=2, ik_chains=2, post_physics_frac=0.05)
C3 = SynthConfig("C3", 3, 1_000_000, 1_000, 200, 300)
C4 = replace(C1, name="C4", config_id=4)
C5 = replace(C3, name="C5", config_id=5, n_frames=10_000)
# Small cases the CPU oracles finish in well under a second.
TINY = SynthConfig("tiny", 10, 2_000, 40, 8, 60)
TINY_FULL = SynthConfig("tiny_full", 11, 3_000, 64, 10, 90, sdef_qdef=True, n_uv_morphs=2, n_group_morphs=2,
                        n_bone_morphs=2, ik_chains=2, post_physics_frac=0.1, stress=True)
SMALL = SynthConfig("small", 12, 20_000, 120, 20, 120, sdef_qdef=True, n_uv_morphs=2, n_group_morphs=2,
                    n_bone_morphs=2, ik_chains=2, post_physics_frac=0.05)
CONFIGS = {c.name: c for c in (C1, C2, C3, C4, C5, TINY, TINY_FULL, SMALL)}


def _rng(cfg: SynthConfig, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([SEED_BASE + cfg.config_id, stream]))


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def make_model(cfg: SynthConfig) -> dict:
    """Flat PMX-shaped arrays for `capi.model_desc`."""
    rng = _rng(cfg, 0)
    nb, nv = cfg.n_bones, cfg.n_vertices
    n_chain = 4 * cfg.ik_chains
    if nb < 1 + n_chain + 2:
        raise ValueError("too few bones for the requested IK chains")

    # ---------------------------------------------------------------- bones
    pos = np.empty((nb, 3), np.float32)
    pos[:, 0] = _f32(rng.uniform(-8, 8, nb))
    pos[:, 1] = _f32(rng.uniform(0, 20, nb))
    pos[:, 2] = _f32(rng.uniform(-3, 3, nb))
    pos[0] = 0
    chain_coin = rng.random(nb) < 0.5
    uni_parent = (rng.random(nb) * np.arange(nb)).astype(np.int64)
    parent = np.where(chain_coin, np.arange(nb) - 1, uni_parent).astype(np.int32)
    parent[0] = -1
    level = np.zeros(nb, np.int32)
    flags = np.zeros(nb, np.uint16)
    append_parent = np.full(nb, -1, np.int32)
    append_ratio = np.zeros(nb, np.float32)
    is_ar = rng.random(nb) < 0.05
    is_at = rng.random(nb) < 0.01
    ratio_pick = rng.integers(0, 3, nb)
    ap_pick = (rng.random(nb) * np.arange(nb)).astype(np.int32)
    special = np.zeros(nb, bool)
    special[: 1 + n_chain] = True           # root and IK chain bones keep plain settings
    is_ar &= ~special
    is_at &= ~special
    has_append = is_ar | is_at
    flags[is_ar] |= capi.BONE_APPEND_ROTATE
    flags[is_at] |= capi.BONE_APPEND_TRANSLATE
    append_parent[has_append] = ap_pick[has_append]
    append_ratio[has_append] = _f32([0.5, 1.0, -0.5])[ratio_pick[has_append]]
    level[has_append] = 1
    if cfg.post_physics_frac > 0:
        pp = (rng.random(nb) < cfg.post_physics_frac) & ~special
        flags[pp] |= capi.BONE_POST_PHYSICS
    if cfg.stress:
        # children that sort before their parent (lower level than the parent), an append bone whose
        # append parent sorts later, and an append flag whose parent index is out of range (dropped,
        # L/motion/poser_impl.inl:51-57)
        cand = np.flatnonzero(~special & ~has_append)
        if cand.size >= 6:
            pick = rng.choice(cand, 6, replace=False)
            level[pick[0]] = 2
            level[pick[1]] = 3
            flags[pick[2]] |= capi.BONE_APPEND_ROTATE
            append_parent[pick[2]] = min(nb - 1, pick[2] + 3)
            append_ratio[pick[2]] = 0.75
            flags[pick[3]] |= capi.BONE_APPEND_TRANSLATE | capi.BONE_APPEND_ROTATE
            append_parent[pick[3]] = nb + 5
            append_ratio[pick[3]] = 1.0
            parent[pick[4]] = -1            # a second root
            flags[pick[5]] |= capi.BONE_POST_PHYSICS

    ik_target = np.full(nb, -1, np.int32)
    ik_iter = np.zeros(nb, np.int32)
    ik_angle = np.zeros(nb, np.float32)
    ik_begin = np.zeros(nb, np.uint32)
    ik_count = np.zeros(nb, np.uint32)
    l_bone, l_has, l_lo, l_hi = [], [], [], []
    for c in range(cfg.ik_chains):
        hip, knee, ankle, ikb = 1 + 4 * c, 2 + 4 * c, 3 + 4 * c, 4 + 4 * c
        x = np.float32(1.0 if c % 2 == 0 else -1.0) * np.float32(1 + c // 2)
        pos[hip] = (x, 10.0, 0.0)
        pos[knee] = (x, 6.0, -0.2)
        pos[ankle] = (x, 1.0, 0.0)
        pos[ikb] = (x, 1.0, 0.0)
        parent[hip], parent[knee], parent[ankle], parent[ikb] = 0, hip, knee, 0
        flags[ikb] |= capi.BONE_HAS_IK
        ik_target[ikb] = ankle
        ik_iter[ikb] = 40
        ik_angle[ikb] = 2.0
        ik_begin[ikb] = len(l_bone)
        ik_count[ikb] = 2
        # tip-most first: knee (limited to X in [-pi, -0.0087]) then hip (free)
        l_bone += [knee, hip]
        l_has += [1, 0]
        l_lo += [(-PI_F, 0, 0), (0, 0, 0)]
        l_hi += [(np.float32(-0.0087266), 0, 0), (0, 0, 0)]
        if cfg.stress and c == 1:
            # second chain: 3-axis limited knee inside (-pi/2, pi/2) -> ZXY order, FIX_NONE,
            # and an odd iteration count so that ikt = iters/2 truncates
            l_lo[-2] = (np.float32(-1.2), np.float32(-0.3), np.float32(-0.2))
            l_hi[-2] = (np.float32(0.1), np.float32(0.3), np.float32(0.25))
            ik_iter[ikb] = 15
            ik_angle[ikb] = 0.6

    # ---------------------------------------------------------------- vertices
    vpos = np.empty((nv, 3), np.float32)
    vpos[:, 0] = _f32(rng.uniform(-8, 8, nv))
    vpos[:, 1] = _f32(rng.uniform(0, 20, nv))
    vpos[:, 2] = _f32(rng.uniform(-3, 3, nv))
    n = _f32(rng.normal(size=(nv, 3)))
    n /= np.maximum(np.sqrt((n * n).sum(1, keepdims=True)), np.float32(1e-6))
    vnrm = _f32(n)
    uv = _f32(rng.random((nv, 2)))
    if cfg.binding in ("coherent", "wide"):
        half = 3 if cfg.binding == "coherent" else 24
        centre = (np.arange(nv, dtype=np.int64) * nb) // nv
        bid = centre[:, None] + rng.integers(-half, half + 1, (nv, 4))
        bid = np.clip(bid, 0, nb - 1).astype(np.int32)
    else:
        bid = rng.integers(0, nb, (nv, 4)).astype(np.int32)
    u = rng.random(nv)
    stype = np.where(u < 0.2, capi.SKIN_BDEF1, np.where(u < 0.7, capi.SKIN_BDEF2, capi.SKIN_BDEF4)).astype(np.uint8)
    w = np.zeros((nv, 4), np.float32)
    w2 = _f32(rng.uniform(0.05, 0.95, nv))
    w4 = _f32(rng.random((nv, 4))) + np.float32(1e-3)
    w4 = _f32(w4 / w4.sum(1, keepdims=True, dtype=np.float32))
    exact = rng.random(nv) < 0.001
    exact_val = _f32(rng.integers(0, 2, nv))
    w2 = np.where(exact, exact_val, w2)
    is1, is2, is4 = stype == capi.SKIN_BDEF1, stype == capi.SKIN_BDEF2, stype == capi.SKIN_BDEF4
    w[is1, 0] = 1.0
    w[is2, 0] = w2[is2]
    w[is4] = w4[is4]
    if cfg.stress:
        # Lerp shortcut thresholds (L/util/math_impl.inl:1246-1250): w just below 1e-7 and just above 1-1e-7
        idx = np.flatnonzero(is2)[:8]
        vals = _f32([5e-8, 1.5e-7, 0.99999994, 0.9999998, 1e-7, 0.99999988, 0.0, 1.0])
        w[idx, 0] = vals[: idx.size]
    sdef_c = sdef_r0 = sdef_r1 = None
    if cfg.sdef_qdef:
        u2 = rng.random(nv)
        to_sdef = is2 & (u2 < 0.15)
        to_qdef = is4 & (u2 < 0.10)
        # parent/child pair: bone 1 of the pair becomes the parent of bone 0 (where it has one)
        p_of = parent[bid[:, 0]]
        related = to_sdef & (p_of >= 0) & (rng.random(nv) < 0.9)
        bid[related, 1] = p_of[related]
        stype[to_sdef] = capi.SKIN_SDEF
        stype[to_qdef] = capi.SKIN_QDEF
        sdef_c = np.zeros((nv, 3), np.float32)
        sdef_r0 = np.zeros((nv, 3), np.float32)
        sdef_r1 = np.zeros((nv, 3), np.float32)
        k = int(to_sdef.sum())
        sdef_c[to_sdef] = vpos[to_sdef] + _f32(rng.uniform(-0.2, 0.2, (k, 3)))
        sdef_r0[to_sdef] = sdef_c[to_sdef] + _f32(rng.uniform(-0.5, 0.5, (k, 3)))
        sdef_r1[to_sdef] = sdef_c[to_sdef] + _f32(rng.uniform(-0.5, 0.5, (k, 3)))

    # ---------------------------------------------------------------- morphs
    run = max(1, int(cfg.morph_run_frac * nv))
    sc = int(cfg.morph_scatter_frac * nv)
    per = run + sc
    nvm = cfg.n_vertex_morphs
    mtype, mbegin, mcount = [], [], []
    ventries = np.zeros(nvm * per, capi.VERTEX_MORPH_ENTRY)
    starts = rng.integers(0, max(1, nv - run + 1), nvm)
    scat = rng.integers(0, nv, (nvm, sc)) if sc else np.zeros((nvm, 0), np.int64)
    offs = _f32(rng.uniform(-0.5, 0.5, (nvm, per, 3)))
    for m in range(nvm):
        vid = np.concatenate([np.arange(starts[m], starts[m] + run) % nv, scat[m]])
        ventries["vertex"][m * per:(m + 1) * per] = vid
        ventries["offset"][m * per:(m + 1) * per] = offs[m]
        mtype.append(capi.MORPH_VERTEX)
        mbegin.append(m * per)
        mcount.append(per)
    uventries = np.zeros(cfg.n_uv_morphs * per, capi.UV_MORPH_ENTRY)
    for m in range(cfg.n_uv_morphs):
        s0 = int(rng.integers(0, max(1, nv - run + 1)))
        vid = np.concatenate([np.arange(s0, s0 + run) % nv, rng.integers(0, nv, sc)])
        uventries["vertex"][m * per:(m + 1) * per] = vid
        uventries["offset"][m * per:(m + 1) * per] = _f32(rng.uniform(-0.1, 0.1, (per, 4)))
        mtype.append(capi.MORPH_UV)
        mbegin.append(m * per)
        mcount.append(per)
    bentries = np.zeros(cfg.n_bone_morphs * 3, capi.BONE_MORPH_ENTRY)
    for m in range(cfg.n_bone_morphs):
        for j in range(3):
            i = m * 3 + j
            bentries["bone"][i] = 1 if (cfg.stress and j == 0 and n_chain) else int(rng.integers(1 + n_chain, nb))
            bentries["translation"][i] = _f32(rng.uniform(-0.3, 0.3, 3))
            q = np.append(_f32(rng.uniform(-0.3, 0.3, 3)), np.float32(1.0))
            q = _f32(q / np.sqrt((q * q).sum(dtype=np.float32)))
            if cfg.stress and j == 1:
                q = -q                                   # SLerp flip branch (math_impl.inl:1314-1317)
            bentries["rotation"][i] = q
        mtype.append(capi.MORPH_BONE)
        mbegin.append(m * 3)
        mcount.append(3)
    gentries = np.zeros(cfg.n_group_morphs * 3, capi.GROUP_MORPH_ENTRY)
    n_before_groups = len(mtype)
    for m in range(cfg.n_group_morphs):
        # children: vertex / bone morphs defined above; in stress mode the second group also contains
        # the first group (group-in-group recursion, poser_impl.inl:334-339)
        kids = rng.choice(n_before_groups, 3, replace=False)
        if cfg.stress and m == 1:
            kids[2] = n_before_groups       # index of group morph 0
        gentries["morph"][m * 3:(m + 1) * 3] = kids
        gentries["rate"][m * 3:(m + 1) * 3] = _f32([0.5, 1.0, 0.25])
        mtype.append(capi.MORPH_GROUP)
        mbegin.append(m * 3)
        mcount.append(3)
    if cfg.stress:
        mtype.append(capi.MORPH_MATERIAL)   # ignored type with no entries
        mbegin.append(0)
        mcount.append(0)
    nm = len(mtype)

    out = dict(
        n_vertices=nv, position=vpos, normal=vnrm, uv=uv, skin_type=stype, bone_id=bid, weight=w,
        n_bones=nb, bone_position=pos, bone_parent=parent, bone_transform_level=level, bone_flags=flags,
        bone_append_parent=append_parent, bone_append_ratio=append_ratio,
        ik_target=ik_target, ik_iterations=ik_iter, ik_angle_limit=ik_angle, ik_link_begin=ik_begin,
        ik_link_count=ik_count,
        n_ik_links=len(l_bone), ik_link_bone=np.asarray(l_bone, np.int32),
        ik_link_has_limit=np.asarray(l_has, np.uint8),
        ik_link_lo=_f32(l_lo).reshape(-1, 3), ik_link_hi=_f32(l_hi).reshape(-1, 3),
        n_morphs=nm, morph_type=np.asarray(mtype, np.uint8), morph_entry_begin=np.asarray(mbegin, np.uint32),
        morph_entry_count=np.asarray(mcount, np.uint32),
        n_vertex_morph_entries=ventries.size, vertex_morph_entries=ventries,
        n_uv_morph_entries=uventries.size, uv_morph_entries=uventries,
        n_bone_morph_entries=bentries.size, bone_morph_entries=bentries,
        n_group_morph_entries=gentries.size, group_morph_entries=gentries,
    )
    if sdef_c is not None:
        out.update(sdef_c=sdef_c, sdef_r0=sdef_r0, sdef_r1=sdef_r1)
    return out


def make_motion(cfg: SynthConfig, model: dict, instance: int = 0, n_frames: int | None = None) -> dict:
    """Flat VMD-shaped arrays for `capi.anim_desc`; `instance` selects an independent clip (crowd)."""
    rng = _rng(cfg, 1000 + instance)
    nb, nm = model["n_bones"], model["n_morphs"]
    nf = cfg.n_frames if n_frames is None else n_frames
    flags = model["bone_flags"]
    phase_on = instance > 0

    t_bone, t_begin, t_count = [], [], []
    key_chunks = []
    total = 0
    moving = rng.random(nb) < 0.10
    for b in range(nb):
        if cfg.stress and b % 13 == 7:
            continue                         # bone without a track stays at the ResetPosing identity
        step = 5 + (b % 11)
        phase = int(rng.integers(0, step)) if phase_on else 0
        frames = np.arange(phase, nf + 1, step, dtype=np.uint32)
        if cfg.stress and b % 13 == 3:
            frames = frames[:0]              # registered but empty track -> identity (motion_impl.inl:260-264)
        k = frames.size
        keys = np.zeros(k, capi.BONE_KEY)
        keys["frame"] = frames
        q = np.concatenate([_f32(rng.uniform(-0.35, 0.35, (k, 3))), np.ones((k, 1), np.float32)], 1)
        q = _f32(q / np.sqrt((q * q).sum(1, keepdims=True, dtype=np.float32)))
        if cfg.stress:
            q[rng.random(k) < 0.3] *= np.float32(-1)      # NLerp dot<0 branch (math_impl.inl:1271-1275)
        keys["rotation"] = q
        is_ik = bool(flags[b] & capi.BONE_HAS_IK)
        if is_ik:
            keys["translation"] = _f32(rng.uniform(-2, 2, (k, 3)) / np.sqrt(3.0))
        elif moving[b]:
            keys["translation"] = _f32(rng.uniform(-0.5, 0.5, (k, 3)))
        interp = rng.integers(0, 128, (k, 4, 4)).astype(np.int8)
        lin = rng.random((k, 4)) < 0.25
        interp[lin] = np.asarray([20, 20, 107, 107], np.int8)
        keys["interp"] = interp
        if cfg.stress and k > 3:
            # unsorted input with a duplicated frame: the later record wins (std::map::operator[])
            dup = keys[2:3].copy()
            dup["rotation"] = keys["rotation"][1]
            keys = np.concatenate([keys[:1], keys[3:], keys[1:3], dup])
        t_bone.append(b)
        t_begin.append(total)
        t_count.append(keys.size)
        total += keys.size
        key_chunks.append(keys)
    bone_keys = np.concatenate(key_chunks) if key_chunks else np.zeros(0, capi.BONE_KEY)

    m_morph, m_begin, m_count, mk_chunks = [], [], [], []
    mtotal = 0
    for m in range(nm):
        phase = int(rng.integers(0, 10)) if phase_on else 0
        frames = np.arange(phase, nf + 1, 10, dtype=np.uint32)
        k = frames.size
        keys = np.zeros(k, capi.MORPH_KEY)
        keys["frame"] = frames
        wgt = _f32(rng.random(k))
        wgt[rng.random(k) < 0.2] = 0
        if cfg.stress:
            neg = rng.random(k) < 0.15
            wgt[neg] = -wgt[neg] - np.float32(0.1)        # negative rates are skipped (poser_impl.inl:329-331)
            tiny = rng.random(k) < 0.05
            wgt[tiny] = np.float32(5e-8)                  # below the 1e-7 skip threshold
        keys["weight"] = wgt
        m_morph.append(m)
        m_begin.append(mtotal)
        m_count.append(k)
        mtotal += k
        mk_chunks.append(keys)
    morph_keys = np.concatenate(mk_chunks) if mk_chunks else np.zeros(0, capi.MORPH_KEY)

    return dict(
        n_bone_tracks=len(t_bone), bone_track_bone=np.asarray(t_bone, np.int32),
        bone_track_key_begin=np.asarray(t_begin, np.uint32), bone_track_key_count=np.asarray(t_count, np.uint32),
        n_bone_keys=bone_keys.size, bone_keys=bone_keys,
        n_morph_tracks=len(m_morph), morph_track_morph=np.asarray(m_morph, np.int32),
        morph_track_key_begin=np.asarray(m_begin, np.uint32), morph_track_key_count=np.asarray(m_count, np.uint32),
        n_morph_keys=morph_keys.size, morph_keys=morph_keys,
    )


def add_material_morphs(model: dict, n_materials: int = 5, n_morphs: int = 4, entries_per_morph: int = 3,
                        seed: int = 0x4D4D4D41, in_group: bool = True) -> dict:
    """Copy of `model` with `n_materials` materials and `n_morphs` material morphs appended after the existing morphs
    (existing indices and random draws are untouched), plus optionally one group morph that drives the first two
    of them.  Every third entry addresses all materials (material = -1, pmx_reader_impl.inl:327-334); methods
    alternate between multiply and add (model.inl:396-399)."""
    rng = np.random.default_rng(seed)
    out = dict(model)
    mtype, mbegin, mcount = (list(model[k]) for k in ("morph_type", "morph_entry_begin", "morph_entry_count"))
    first = len(mtype)
    ent = np.zeros(n_morphs * entries_per_morph, capi.MATERIAL_MORPH_ENTRY)
    ent["material"] = rng.integers(0, n_materials, ent.size)
    ent["material"][2::3] = -1
    ent["method"] = np.arange(ent.size) % 2
    ent["value"] = _f32(rng.uniform(0.0, 2.0, (ent.size, capi.MATERIAL_FIELDS)))
    for m in range(n_morphs):
        mtype.append(capi.MORPH_MATERIAL)
        mbegin.append(m * entries_per_morph)
        mcount.append(entries_per_morph)
    if in_group and n_morphs >= 2:
        g = np.zeros(2, capi.GROUP_MORPH_ENTRY)
        g["morph"] = [first, first + 1]
        g["rate"] = _f32([0.5, 0.75])
        old = np.asarray(model["group_morph_entries"], capi.GROUP_MORPH_ENTRY)
        mtype.append(capi.MORPH_GROUP)
        mbegin.append(old.size)
        mcount.append(2)
        gall = np.concatenate([old, g])
        out.update(n_group_morph_entries=gall.size, group_morph_entries=gall)
    out.update(n_morphs=len(mtype), morph_type=np.asarray(mtype, np.uint8),
               morph_entry_begin=np.asarray(mbegin, np.uint32), morph_entry_count=np.asarray(mcount, np.uint32),
               n_materials=n_materials, n_material_morph_entries=ent.size, material_morph_entries=ent)
    return out


def mean_morph_entries_per_vertex(model: dict) -> float:
    """e of SURVEY 8d: vertex-morph entries reachable through application slots, per vertex."""
    return float(model["n_vertex_morph_entries"]) / max(1, model["n_vertices"])


def make_ik_zoo(seed: int = 77, n_frames: int = 40, chains=None):
    """A small rig that exercises every CCD IK branch of L/motion/poser_impl.inl:168-310: chains of 2-4 links,
    unlimited links, FIX_X / FIX_Y / FIX_Z / FIX_ALL links, the three Euler orders (ZXY, XYZ, YZX), swapped lo / hi
    limits, odd iteration counts, an iteration count above the 256 cap, and an IK bone that sorts before its links.
    Returns (model, motion) in the flat layout of `capi.model_desc` / `capi.anim_desc`."""
    rng = np.random.Generator(np.random.PCG64([SEED_BASE + 99, seed]))
    PI = float(np.pi)
    zoo = [
        # (n_links, iterations, angle_limit, [per link (tip-most first): None | (lo, hi)][, IK bone sorts first])
        (2, 40, 2.0, [((-PI, 0, 0), (-0.0087, 0, 0)), None]),                           # FIX_X, XYZ order (classic knee)
        (3, 15, 0.6, [((0, -1.0, 0), (0, 1.2, 0)), None, ((-0.2, -0.3, -0.4), (0.3, 0.2, 0.5))]),   # FIX_Y ; ZXY
        (3, 33, 1.0, [((0, 0, -2.5), (0, 0, 0.4)), ((0, 0, 0), (0, 0, 0)), None]),      # FIX_Z ; FIX_ALL (skipped)
        (4, 300, 0.35, [((-2.0, -2.0, -0.5), (2.0, 2.0, 0.5)), None, ((0.4, -0.2, -0.1), (-0.9, 0.3, 0.2)), None]),  # YZX ; swapped lo/hi
        (2, 7, 3.5, [((-2.0, -1.0, -3.0), (2.0, 1.0, 3.0)), ((-1.0, -2.5, -1.0), (1.0, 2.5, 1.0))]),  # XYZ ; ZXY
    ]
    if chains is None:
        chains = zoo                              # custom chain lists: tools/gpu_fuzz.py
    pos, parent, level, flags = [(0.0, 0.0, 0.0)], [-1], [0], [0]
    ik_target, ik_iter, ik_angle, ik_begin, ik_count = [-1], [0], [0.0], [0], [0]
    l_bone, l_has, l_lo, l_hi = [], [], [], []
    bind = []

    def add_bone(p, par, lvl=0, fl=0):
        pos.append(tuple(float(x) for x in p)); parent.append(par); level.append(lvl); flags.append(fl)
        ik_target.append(-1); ik_iter.append(0); ik_angle.append(0.0); ik_begin.append(0); ik_count.append(0)
        return len(pos) - 1

    for c, spec in enumerate(chains):
        nl, iters, angle, lims = spec[:4]
        x = -6.0 + 3.0 * c
        ik_first = spec[4] if len(spec) > 4 else (c == 3)   # this IK bone is created (and therefore sorts) before its links
        ikb = add_bone((x, 1.0, 0.3), 0, 0, capi.BONE_HAS_IK) if ik_first else None
        links = []
        par = 0
        for j in range(nl):                       # root-most first
            y = 12.0 - 11.0 * j / nl
            b = add_bone((x + 0.1 * j, y, 0.15 * ((-1) ** j)), par)
            links.append(b)
            bind.append(b)
            par = b
        tgt = add_bone((x, 1.0, 0.0), par)
        bind.append(tgt)
        if ikb is None:
            ikb = add_bone((x, 1.0, 0.3), 0, 0, capi.BONE_HAS_IK)
        ik_target[ikb] = tgt
        ik_iter[ikb] = iters
        ik_angle[ikb] = angle
        ik_begin[ikb] = len(l_bone)
        ik_count[ikb] = nl
        for j, lim in enumerate(lims):            # tip-most first
            l_bone.append(links[nl - 1 - j])
            l_has.append(0 if lim is None else 1)
            l_lo.append((0, 0, 0) if lim is None else lim[0])
            l_hi.append((0, 0, 0) if lim is None else lim[1])
    nb = len(pos)
    nv = 40 * len(bind)
    vpos = np.empty((nv, 3), np.float32)
    bid = np.zeros((nv, 4), np.int32)
    w = np.zeros((nv, 4), np.float32)
    stype = np.zeros(nv, np.uint8)
    bpos = _f32(pos)
    for i in range(nv):
        b = bind[i // 40]
        vpos[i] = bpos[b] + _f32(rng.uniform(-0.5, 0.5, 3))
        if i % 2 == 0:
            stype[i] = capi.SKIN_BDEF1
            bid[i, 0] = b
            w[i, 0] = 1.0
        else:
            stype[i] = capi.SKIN_BDEF2
            bid[i, 0], bid[i, 1] = b, max(0, parent[b])
            w[i, 0] = np.float32(rng.uniform(0.1, 0.9))
    n = _f32(rng.normal(size=(nv, 3)))
    n /= np.sqrt((n * n).sum(1, keepdims=True))
    model = dict(
        n_vertices=nv, position=vpos, normal=_f32(n), uv=_f32(rng.random((nv, 2))), skin_type=stype, bone_id=bid, weight=w,
        n_bones=nb, bone_position=bpos, bone_parent=np.asarray(parent, np.int32),
        bone_transform_level=np.asarray(level, np.int32), bone_flags=np.asarray(flags, np.uint16),
        bone_append_parent=np.full(nb, -1, np.int32), bone_append_ratio=np.zeros(nb, np.float32),
        ik_target=np.asarray(ik_target, np.int32), ik_iterations=np.asarray(ik_iter, np.int32),
        ik_angle_limit=_f32(ik_angle), ik_link_begin=np.asarray(ik_begin, np.uint32), ik_link_count=np.asarray(ik_count, np.uint32),
        n_ik_links=len(l_bone), ik_link_bone=np.asarray(l_bone, np.int32), ik_link_has_limit=np.asarray(l_has, np.uint8),
        ik_link_lo=_f32(l_lo).reshape(-1, 3), ik_link_hi=_f32(l_hi).reshape(-1, 3),
        n_morphs=0, morph_type=np.zeros(0, np.uint8), morph_entry_begin=np.zeros(0, np.uint32),
        morph_entry_count=np.zeros(0, np.uint32),
        n_vertex_morph_entries=0, vertex_morph_entries=np.zeros(0, capi.VERTEX_MORPH_ENTRY),
        n_uv_morph_entries=0, uv_morph_entries=np.zeros(0, capi.UV_MORPH_ENTRY),
        n_bone_morph_entries=0, bone_morph_entries=np.zeros(0, capi.BONE_MORPH_ENTRY),
        n_group_morph_entries=0, group_morph_entries=np.zeros(0, capi.GROUP_MORPH_ENTRY),
    )
    # motion: every bone keyed every 4 frames; IK bones get translations that pull the targets around
    t_bone, t_begin, t_count, chunks, total = [], [], [], [], 0
    for b in range(nb):
        frames = np.arange(0, n_frames + 1, 4, dtype=np.uint32)
        k = frames.size
        keys = np.zeros(k, capi.BONE_KEY)
        keys["frame"] = frames
        q = np.concatenate([_f32(rng.uniform(-0.3, 0.3, (k, 3))), np.ones((k, 1), np.float32)], 1)
        keys["rotation"] = _f32(q / np.sqrt((q * q).sum(1, keepdims=True, dtype=np.float32)))
        if flags[b] & capi.BONE_HAS_IK:
            keys["translation"] = _f32(rng.uniform(-2.5, 2.5, (k, 3)))
        interp = rng.integers(0, 128, (k, 4, 4)).astype(np.int8)
        keys["interp"] = interp
        t_bone.append(b); t_begin.append(total); t_count.append(k); total += k
        chunks.append(keys)
    bone_keys = np.concatenate(chunks)
    motion = dict(
        n_bone_tracks=len(t_bone), bone_track_bone=np.asarray(t_bone, np.int32),
        bone_track_key_begin=np.asarray(t_begin, np.uint32), bone_track_key_count=np.asarray(t_count, np.uint32),
        n_bone_keys=bone_keys.size, bone_keys=bone_keys,
        n_morph_tracks=0, morph_track_morph=np.zeros(0, np.int32), morph_track_key_begin=np.zeros(0, np.uint32),
        morph_track_key_count=np.zeros(0, np.uint32), n_morph_keys=0, morph_keys=np.zeros(0, capi.MORPH_KEY),
    )
    return model, motion


def make_ik_nested(seed: int = 78, n_frames: int = 40):
    """A rig whose CCD IK solves nest: libmmd re-evaluates a solve's links and target with UpdateBoneTransform
    (L/motion/poser_impl.inl:203-206, :303), which re-enters the IK block when that bone has IK itself.
      * chain 0's TARGET is an IK bone (its own solve runs after every CCD step of the outer one),
      * chain 2's middle LINK is an IK bone (its solve runs once, when the outer solve first re-evaluates its links),
      * chain 4 -> its target -> that solve's target: three levels.
    Built from make_ik_zoo's chains by moving the IK record of a helper chain onto a link / target of another one."""
    PI = float(np.pi)
    chains = [
        (2, 12, 1.5, [((-PI, 0, 0), (-0.0087, 0, 0)), None], False),                     # 0: outer, target gets chain 1's solve
        (1, 6, 2.0, [None], False),                                                     # 1: helper
        (3, 9, 0.8, [None, ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5)), None], False),         # 2: outer, middle link gets chain 3's solve
        (2, 5, 1.0, [((0, -1.0, 0), (0, 1.2, 0)), None], False),                         # 3: helper
        (2, 8, 1.2, [None, None], False),                                               # 4: outer (level 1)
        (1, 4, 2.5, [((-2.0, -1.0, -3.0), (2.0, 1.0, 3.0))], False),                     # 5: helper (level 2), moved onto 4's target
        (2, 3, 0.7, [None, None], False),                                               # 6: helper (level 3), moved onto 5's target
    ]
    model, motion = make_ik_zoo(seed=seed, n_frames=n_frames, chains=chains)
    # bone indices by construction order: per chain its links (root-most first), its target, its IK bone
    first, layout = 1, []
    for spec in chains:
        nl = spec[0]
        layout.append(dict(links=list(range(first, first + nl)), target=first + nl, ik=first + nl + 1))
        first += nl + 2
    flags = model["bone_flags"].copy()

    def move_ik(src_chain: int, dst_bone: int):
        src = layout[src_chain]["ik"]
        for k in ("ik_target", "ik_iterations", "ik_angle_limit", "ik_link_begin", "ik_link_count"):
            a = model[k].copy()
            a[dst_bone] = a[src]
            a[src] = -1 if k == "ik_target" else 0
            model[k] = a
        flags[dst_bone] |= capi.BONE_HAS_IK
        flags[src] &= ~np.uint16(capi.BONE_HAS_IK)

    move_ik(1, layout[0]["target"])
    move_ik(3, layout[2]["links"][1])
    move_ik(5, layout[4]["target"])
    move_ik(6, layout[5]["target"])
    model["bone_flags"] = flags
    return model, motion
