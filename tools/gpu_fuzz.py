"""GPU box: randomized model / motion configurations beyond the seeds the test-suite pins, each compared bit-for-bit
with the CPU oracle (positions, normals, bone matrices, sampled poses, morph rates).
  phase "rig":  varied rigs through the fused update in both output layouts, plus MotionPlayer::SeekTime at random
                times and the step-wise libmmd call sequence (must equal the fused path bit-for-bit)
  phase "ik":   random CCD IK chains (1-5 links, 1-300 iterations, angle limits, per-axis limits with zero ranges /
                swapped bounds / sub-quadrant ranges that select the three Euler orders)
  phase "topo": arbitrary skeleton topology - parents that are later bones or the bone itself, mixed transform levels,
                append parents anywhere (self included), post-physics bones, IK chains over arbitrary bone sets that
                share links and targets
  phase "morph": random morph graphs (nested groups with negative / zero / sub-epsilon rates, bone morphs on any bone,
                vertex morphs with repeated vertices) driven through SetBonePose / SetMorphPose with edge weights
  phase "motion": random key-frame structure (empty / unsorted / repeated / far-away keys, extreme and linear Bezier
                bytes, both quaternion hemispheres) sampled by SeekFrame and SeekTime, incl. far past the clip
  phase "crowd": instances x frames with per-instance clips, range mode with a stride and per-slot frame ids
  phase "ext":   extensions = 1: non-extension vertices bit-exact, SDEF / QDEF / UV / material images vs the fp64 restatement
usage: python tools/gpu_fuzz.py [first_seed] [count] [rig|ik|topo|nest|morph|motion|crowd|ext|all]"""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from dataclasses import replace
import numpy as np
import oracle
from simple_mmd_renderer_b200 import capi, synth
from simple_mmd_renderer_b200.poser import Context, Frames, MmdGpuError, Model, Motion

ctx = None   # created by main()


def same(a, b):
    """Bit-identical.  Frames in which libmmd itself produces non-finite values are not compared: the device keeps
    matrices as 4 x 3 (fourth column implied), libmmd carries a fourth column that turns NaN once an element
    overflows (seen only with a self-parented bone inside an IK chain, which squares its matrix every CCD step),
    so the SET of NaN elements can differ; finite results never do."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if not np.isfinite(b).all():
        return True
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


def check_slot(fr, k, ref):
    ok = same(fr.download(k, capi.STREAM_POSITION), ref["pos"]) and same(fr.download(k, capi.STREAM_NORMAL), ref["nrm"])
    ok &= same(fr.bone_matrices(k), ref["skin"]) and same(fr.bone_poses(k), ref["poses"])
    if ref["rates"].size:
        ok &= same(fr.morph_rates(k), ref["rates"])
    return ok


def rig_case(seed):
    rng = np.random.default_rng(1000 + seed)
    ik = int(rng.integers(0, 3))
    cfg = replace(synth.TINY_FULL, name=f"rand{seed}", config_id=200 + seed,
                  n_bones=int(rng.integers(4 * ik + 12, 400)), n_vertices=int(rng.integers(1, 6000)),
                  n_vertex_morphs=int(rng.integers(0, 40)), n_frames=int(rng.integers(8, 80)),
                  binding=("coherent", "random")[int(rng.integers(0, 2))], ik_chains=ik,
                  n_group_morphs=int(rng.integers(0, 3)), n_bone_morphs=int(rng.integers(0, 3)),
                  n_uv_morphs=int(rng.integers(0, 3)), post_physics_frac=float(rng.choice([0.0, 0.1, 0.4])),
                  stress=bool(rng.integers(0, 2)), morph_run_frac=float(rng.choice([0.018, 0.2])),
                  morph_scatter_frac=float(rng.choice([0.002, 0.05])))
    if cfg.n_group_morphs and cfg.n_vertex_morphs + cfg.n_uv_morphs + cfg.n_bone_morphs < 3:
        cfg = replace(cfg, n_group_morphs=0)
    model = synth.make_model(cfg)
    motion = synth.make_motion(cfg, model)
    orc = oracle.Restatement(model, motion)
    m = Model(ctx, model)
    a = Motion(m, motion)
    n_slots = int(rng.integers(1, 9))
    frames = [int(x) for x in rng.integers(0, cfg.n_frames + 5, n_slots)]
    fr = Frames(m, 1, n_slots)
    fr.update(a, frames)
    fi = Frames(m, 1, n_slots, capi.LAYOUT_INTERLEAVED_SOKOL32)
    fi.update(a, frames)
    ok = True
    for k, f in enumerate(frames):
        ref = orc.run_frame(f)
        ok &= check_slot(fr, k, ref)
        ok &= same(fi.download(k, capi.STREAM_INTERLEAVED), orc.repack_sokol32())
    # step-wise libmmd sequence == fused path
    fused_pos = [fr.download(k, capi.STREAM_POSITION) for k in range(n_slots)]
    fr.reset_posing(); fr.seek_frame(a, frames); fr.pre_physics_posing(); fr.post_physics_posing(); fr.deform()
    for k in range(n_slots):
        ok &= same(fr.download(k, capi.STREAM_POSITION), fused_pos[k])
    # MotionPlayer::SeekTime at arbitrary times (sub-frame sampling)
    times = rng.uniform(0.0, (cfg.n_frames + 3) / 30.0, n_slots)
    fr.reset_posing(); fr.seek_time(a, times); fr.pre_physics_posing(); fr.post_physics_posing(); fr.deform()
    for k, t in enumerate(times):
        ref = orc.run_time(float(t))
        ok &= same(fr.download(k, capi.STREAM_POSITION), ref["pos"]) and same(fr.bone_poses(k), ref["poses"])
        ok &= same(fr.bone_matrices(k), ref["skin"])
    fr.close(); fi.close(); orc.close()
    return ok, str(cfg)


def ik_case(seed):
    rng = np.random.default_rng(5000 + seed)
    PI = float(np.pi)

    def limit():
        lo, hi = [0.0] * 3, [0.0] * 3
        for ax in range(3):
            kind = rng.random()
            if kind < 0.35:
                continue                                        # zero range: a fixed axis (poser_impl.inl:83-91)
            if kind < 0.45:
                lo[ax], hi[ax] = float(rng.uniform(0, 1.5)), float(rng.uniform(-1.5, 0))          # swapped
            elif kind < 0.7:
                lo[ax], hi[ax] = float(rng.uniform(-1.5, 0)), float(rng.uniform(0, 1.5))          # inside (-pi/2, pi/2)
            else:
                lo[ax], hi[ax] = float(rng.uniform(-PI, 0)), float(rng.uniform(0, PI))
            if rng.random() < 0.1:
                lo[ax] = 5e-8                                   # below the 1e-7 "is zero" threshold
        return tuple(lo), tuple(hi)

    chains = []
    for _ in range(int(rng.integers(1, 6))):
        nl = int(rng.integers(1, 6))
        iters = int(rng.choice([1, 2, 3, 7, 15, 40, 41, 300]))
        angle = float(rng.choice([0.05, 0.35, 1.0, 2.0, 3.5, 6.5]))
        lims = [None if rng.random() < 0.4 else limit() for _ in range(nl)]
        chains.append((nl, iters, angle, lims, bool(rng.random() < 0.25)))
    model, motion = synth.make_ik_zoo(seed=1000 + seed, n_frames=24, chains=chains)
    orc = oracle.Restatement(model, motion)
    m = Model(ctx, model)
    a = Motion(m, motion)
    frames = [int(x) for x in rng.integers(0, 28, 6)]
    fr = Frames(m, 1, len(frames))
    fr.update(a, frames)
    ok = True
    for k, f in enumerate(frames):
        ok &= check_slot(fr, k, orc.run_frame(f))
    fr.close(); orc.close()
    return ok, str(chains)


TOPO_MASK = int(os.environ.get("TOPO_MASK", "31"))   # 1 parents, 2 levels, 4 appends, 8 post-physics, 16 IK


def topo_case_inputs(seed, nest=False):
    rng = np.random.default_rng((19000 if nest else 9000) + seed)
    cfg = replace(synth.TINY_FULL, name=f"topo{seed}", config_id=400 + seed, n_bones=int(rng.integers(8, 60)),
                  n_vertices=int(rng.integers(50, 1500)), ik_chains=0, n_frames=20, stress=bool(rng.integers(0, 2)),
                  n_bone_morphs=int(rng.integers(0, 3)), n_group_morphs=0, n_uv_morphs=0, n_vertex_morphs=int(rng.integers(0, 6)))
    model = dict(synth.make_model(cfg))
    nb = int(model["n_bones"])
    parent = model["bone_parent"].copy(); level = model["bone_transform_level"].copy()
    flags = model["bone_flags"].copy(); ap = model["bone_append_parent"].copy(); ratio = model["bone_append_ratio"].copy()
    for b in range(nb):
        if rng.random() < 0.4 and (TOPO_MASK & 1):
            parent[b] = int(rng.integers(-1, nb))            # any bone, itself and later ones included
        if rng.random() < 0.3 and (TOPO_MASK & 2):
            level[b] = int(rng.integers(0, 3))
        if rng.random() < 0.2 and (TOPO_MASK & 4):
            flags[b] |= int(rng.choice([capi.BONE_APPEND_ROTATE, capi.BONE_APPEND_TRANSLATE, capi.BONE_APPEND_ROTATE | capi.BONE_APPEND_TRANSLATE]))
            ap[b] = int(rng.integers(0, nb))
            ratio[b] = float(rng.choice([0.5, 1.0, -0.5, 0.25, 2.0]))
        if rng.random() < 0.1 and (TOPO_MASK & 8):
            flags[b] |= capi.BONE_POST_PHYSICS
    n_ik = (int(rng.integers(2, 6)) if nest else int(rng.integers(0, 4))) if (TOPO_MASK & 16) else 0
    ik_target = np.full(nb, -1, np.int32); ik_iter = np.zeros(nb, np.int32); ik_angle = np.zeros(nb, np.float32)
    ik_begin = np.zeros(nb, np.uint32); ik_count = np.zeros(nb, np.uint32)
    l_bone, l_has, l_lo, l_hi = [], [], [], []
    ik_bones = [int(x) for x in rng.choice(nb, n_ik, replace=False)] if n_ik else []
    others = [b for b in range(nb) if b not in ik_bones]
    for ikb in ik_bones:
        if len(others) < 3: break
        pool = [b for b in range(nb) if b != ikb] if nest else others      # nest: links / targets may be IK bones themselves
        pick = [int(x) for x in rng.choice(pool, min(len(pool), int(rng.integers(2, 6))), replace=False)]
        tgt, links = pick[0], pick[1:]
        flags[ikb] |= capi.BONE_HAS_IK
        ik_target[ikb] = tgt; ik_iter[ikb] = int(rng.choice([1, 3, 8, 20])); ik_angle[ikb] = float(rng.choice([0.3, 1.0, 2.0]))
        ik_begin[ikb] = len(l_bone); ik_count[ikb] = len(links)
        for l in links:
            l_bone.append(l)
            if rng.random() < 0.5:
                l_has.append(0); l_lo.append((0, 0, 0)); l_hi.append((0, 0, 0))
            else:
                l_has.append(1)
                lo = [float(rng.uniform(-2, 0)) if rng.random() < 0.6 else 0.0 for _ in range(3)]
                hi = [float(rng.uniform(0, 2)) if lo[i] != 0.0 else 0.0 for i in range(3)]
                l_lo.append(tuple(lo)); l_hi.append(tuple(hi))
    model.update(bone_parent=parent, bone_transform_level=level, bone_flags=flags, bone_append_parent=ap, bone_append_ratio=ratio,
                 ik_target=ik_target, ik_iterations=ik_iter, ik_angle_limit=ik_angle, ik_link_begin=ik_begin, ik_link_count=ik_count,
                 n_ik_links=len(l_bone), ik_link_bone=np.asarray(l_bone, np.int32), ik_link_has_limit=np.asarray(l_has, np.uint8),
                 ik_link_lo=np.asarray(l_lo, np.float32).reshape(-1, 3), ik_link_hi=np.asarray(l_hi, np.float32).reshape(-1, 3))
    motion = synth.make_motion(cfg, model)
    frames = [int(x) for x in rng.integers(0, 24, 4)]
    return model, motion, frames


def topo_case(seed, nest=False):
    model, motion, frames = topo_case_inputs(seed, nest)
    try:
        m = Model(ctx, model)
    except MmdGpuError as e:
        # solves that reach each other (libmmd recurses forever: the oracle must not be run) or nest deeper than 3 levels
        if nest and e.status in (capi.ERR_BAD_INDEX, capi.ERR_UNSUPPORTED):
            return True, f"nested topology seed {seed}: refused ({e.message})"
        raise
    orc = oracle.Restatement(model, motion)
    a = Motion(m, motion)
    fr = Frames(m, 1, len(frames))
    fr.update(a, frames)
    ok = True
    for k, f in enumerate(frames):
        ok &= check_slot(fr, k, orc.run_frame(f))
    fr.close(); orc.close()
    return ok, f"{'nested ' if nest else ''}topology seed {seed}"


def nest_case(seed):
    """Topology phase with IK bones allowed as links / targets of other solves (poser_impl.inl:203-206, :303)."""
    return topo_case(seed, nest=True)


def morph_case_inputs(seed):
    """Random morph graph on a small rig: vertex morphs with duplicate vertices, bone morphs on any bone (IK links
    and IK bones included), group morphs nested up to four deep with negative / zero / below-epsilon / > 1 rates,
    UV and material morphs in between; driven through SetBonePose / SetMorphPose with weights that include
    negatives, 5e-8, 1e-7 and values above 1."""
    rng = np.random.default_rng(7000 + seed)
    cfg = replace(synth.TINY_FULL, name=f"morph{seed}", config_id=600 + seed, n_bones=int(rng.integers(14, 50)),
                  n_vertices=int(rng.integers(30, 1200)), ik_chains=int(rng.integers(0, 3)), n_frames=10,
                  stress=False, n_bone_morphs=0, n_group_morphs=0, n_uv_morphs=0, n_vertex_morphs=0)
    model = dict(synth.make_model(cfg))
    nv, nb = int(model["n_vertices"]), int(model["n_bones"])
    nm = int(rng.integers(1, 24))
    order = rng.permutation(nm)                 # group children must come later in this order: a DAG, no cycles
    rank = np.empty(nm, np.int64); rank[order] = np.arange(nm)
    mtype, mbegin, mcount = [], [], []
    ve, be, ge, ue = [], [], [], []
    for m in range(nm):
        later = [int(x) for x in order[rank[m] + 1:]]
        kind = rng.random()
        if kind < 0.25 and later:
            kids = [int(x) for x in rng.choice(later, min(len(later), int(rng.integers(1, 5))), replace=True)]
            mtype.append(capi.MORPH_GROUP); mbegin.append(len(ge)); mcount.append(len(kids))
            for k in kids:
                ge.append((k, float(rng.choice([-0.5, 0.0, 1e-8, 0.25, 0.5, 1.0, 2.0]))))
        elif kind < 0.45:
            n = int(rng.integers(1, 5))
            mtype.append(capi.MORPH_BONE); mbegin.append(len(be)); mcount.append(n)
            for _ in range(n):
                q = rng.normal(size=4); q /= np.linalg.norm(q)
                be.append((int(rng.integers(0, nb)), tuple(rng.uniform(-1, 1, 3)), tuple(q)))
        elif kind < 0.55:
            n = int(rng.integers(1, 20))
            mtype.append(capi.MORPH_UV); mbegin.append(len(ue)); mcount.append(n)
            for _ in range(n):
                ue.append((int(rng.integers(0, nv)), tuple(rng.uniform(-1, 1, 4))))
        else:
            n = int(rng.integers(1, max(2, nv // 2)))
            mtype.append(capi.MORPH_VERTEX); mbegin.append(len(ve)); mcount.append(n)
            for v in rng.integers(0, nv, n):     # with repeats
                ve.append((int(v), tuple(rng.uniform(-0.5, 0.5, 3))))
    def pool(items, dtype, fields):
        a = np.zeros(len(items), dtype)
        for i, it in enumerate(items):
            for f, x in zip(fields, it):
                a[f][i] = x
        return a
    model.update(
        n_morphs=nm, morph_type=np.asarray(mtype, np.uint8), morph_entry_begin=np.asarray(mbegin, np.uint32),
        morph_entry_count=np.asarray(mcount, np.uint32),
        vertex_morph_entries=pool(ve, capi.VERTEX_MORPH_ENTRY, ("vertex", "offset")), n_vertex_morph_entries=len(ve),
        bone_morph_entries=pool(be, capi.BONE_MORPH_ENTRY, ("bone", "translation", "rotation")), n_bone_morph_entries=len(be),
        group_morph_entries=pool(ge, capi.GROUP_MORPH_ENTRY, ("morph", "rate")), n_group_morph_entries=len(ge),
        uv_morph_entries=pool(ue, capi.UV_MORPH_ENTRY, ("vertex", "offset")), n_uv_morph_entries=len(ue))
    poses = []
    for _ in range(3):
        nbp = int(rng.integers(0, nb))
        bones = rng.choice(nb, nbp, replace=False).astype(np.int32)
        q = rng.normal(size=(nbp, 4)); q[:, 3] += 2.0; q /= np.linalg.norm(q, axis=1, keepdims=True)
        p7 = np.concatenate([rng.uniform(-1, 1, (nbp, 3)), q], 1).astype(np.float32)
        morphs = np.arange(nm, dtype=np.int32)
        w = rng.choice([-0.3, 0.0, 5e-8, 1e-7, 1.5e-7, 0.2, 0.5, 1.0, 1.7], nm).astype(np.float32)
        poses.append((bones, p7, morphs, w))
    return model, poses


def morph_case(seed):
    model, poses = morph_case_inputs(seed)
    orc = oracle.Restatement(model, None)
    m = Model(ctx, model)
    fr = Frames(m, 1, len(poses))
    fr.reset_posing()
    for slot, (bones, p7, morphs, w) in enumerate(poses):
        for b, p in zip(bones, p7):
            fr.set_bone_pose(slot, int(b), p[:3], p[3:])
        for i, x in zip(morphs, w):
            fr.set_morph_pose(slot, int(i), float(x))
    fr.pre_physics_posing(); fr.post_physics_posing(); fr.deform()
    ok = True
    for slot, (bones, p7, morphs, w) in enumerate(poses):
        ref = orc.run_manual(bones, p7, morphs, w)
        ok &= same(fr.download(slot, capi.STREAM_POSITION), ref["pos"]) and same(fr.download(slot, capi.STREAM_NORMAL), ref["nrm"])
        ok &= same(fr.bone_matrices(slot), ref["skin"])
    fr.close(); orc.close()
    return ok, f"morph graph seed {seed}"


def motion_case_inputs(seed):
    """Random key-frame structure: 0-7 keys per track (0 = registered but empty), unsorted, repeated frames (last one
    wins), frames far beyond the others, Bezier bytes over the whole 0..127 range plus exact-linear quadruples,
    rotations in both hemispheres (NLerp's dot < 0 branch) and repeated poses; morph weights with negatives."""
    rng = np.random.default_rng(11000 + seed)
    cfg = replace(synth.TINY, name=f"motion{seed}", config_id=800 + seed, n_bones=int(rng.integers(6, 40)),
                  n_vertices=int(rng.integers(20, 600)), n_vertex_morphs=int(rng.integers(0, 8)), n_frames=30)
    model = synth.make_model(cfg)
    nb, nm = int(model["n_bones"]), int(model["n_morphs"])
    t_bone, t_begin, t_count, keys = [], [], [], []
    for b in range(nb):
        if rng.random() < 0.2:
            continue
        n = int(rng.integers(0, 8))
        fr = rng.integers(0, 40, n)
        if n and rng.random() < 0.3:
            fr[rng.integers(0, n)] = int(rng.choice([0, 1000, 100000]))
        if n > 1 and rng.random() < 0.4:
            fr[1] = fr[0]
        k = np.zeros(n, capi.BONE_KEY)
        k["frame"] = fr
        k["translation"] = rng.uniform(-2, 2, (n, 3)) * (rng.random((n, 1)) < 0.5)
        q = rng.normal(size=(n, 4)); q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-9)
        if n > 1 and rng.random() < 0.3:
            q[1] = q[0]
        k["rotation"] = q
        ip = rng.integers(0, 128, (n, 4, 4))
        lin = rng.random((n, 4)) < 0.3
        ip[lin] = [20, 20, 107, 107]
        ext = rng.random((n, 4)) < 0.1
        ip[ext] = rng.choice([0, 127], (int(ext.sum()), 4))
        k["interp"] = ip
        t_bone.append(b); t_begin.append(sum(x.size for x in keys)); t_count.append(n); keys.append(k)
    bone_keys = np.concatenate(keys) if keys else np.zeros(0, capi.BONE_KEY)
    m_morph, m_begin, m_count, mkeys = [], [], [], []
    for m in range(nm):
        if rng.random() < 0.2:
            continue
        n = int(rng.integers(0, 6))
        k = np.zeros(n, capi.MORPH_KEY)
        k["frame"] = rng.integers(0, 40, n)
        k["weight"] = rng.choice([-0.5, 0.0, 5e-8, 0.3, 1.0, 1.5], n)
        m_morph.append(m); m_begin.append(sum(x.size for x in mkeys)); m_count.append(n); mkeys.append(k)
    morph_keys = np.concatenate(mkeys) if mkeys else np.zeros(0, capi.MORPH_KEY)
    motion = dict(n_bone_tracks=len(t_bone), bone_track_bone=np.asarray(t_bone, np.int32),
                  bone_track_key_begin=np.asarray(t_begin, np.uint32), bone_track_key_count=np.asarray(t_count, np.uint32),
                  n_bone_keys=bone_keys.size, bone_keys=bone_keys,
                  n_morph_tracks=len(m_morph), morph_track_morph=np.asarray(m_morph, np.int32),
                  morph_track_key_begin=np.asarray(m_begin, np.uint32), morph_track_key_count=np.asarray(m_count, np.uint32),
                  n_morph_keys=morph_keys.size, morph_keys=morph_keys)
    frames = [int(x) for x in rng.integers(0, 45, 5)] + [int(rng.choice([999, 1000, 1001, 200000]))]
    times = [float(x) for x in rng.uniform(0, 1.5, 5)] + [float(rng.choice([0.0, 1.0 / 30.0, 33.34, 7000.0]))]
    return model, motion, frames, times


def motion_case(seed):
    model, motion, frames, times = motion_case_inputs(seed)
    orc = oracle.Restatement(model, motion)
    m = Model(ctx, model)
    a = Motion(m, motion)
    fr = Frames(m, 1, len(frames))
    fr.update(a, frames)
    ok = True
    for k, f in enumerate(frames):
        ok &= check_slot(fr, k, orc.run_frame(f))
    fr.reset_posing(); fr.seek_time(a, times); fr.pre_physics_posing(); fr.post_physics_posing(); fr.deform()
    for k, t in enumerate(times):
        ref = orc.run_time(t)
        ok &= same(fr.download(k, capi.STREAM_POSITION), ref["pos"]) and same(fr.bone_poses(k), ref["poses"])
        ok &= same(fr.bone_matrices(k), ref["skin"])
        if ref["rates"].size:
            ok &= same(fr.morph_rates(k), ref["rates"])
    fr.close(); orc.close()
    return ok, f"motion structure seed {seed}"


def crowd_case(seed):
    """Instances x frames: per-instance clips, range mode with a stride, then per-slot frame ids on the same object."""
    rng = np.random.default_rng(13000 + seed)
    cfg = replace(synth.TINY_FULL, name=f"crowd{seed}", config_id=900 + seed, n_bones=int(rng.integers(12, 60)),
                  n_vertices=int(rng.integers(1, 1500)), n_frames=int(rng.integers(10, 40)), ik_chains=int(rng.integers(0, 2)),
                  stress=bool(rng.integers(0, 2)))
    model = synth.make_model(cfg)
    ni, nf, stride = int(rng.integers(1, 6)), int(rng.integers(1, 8)), int(rng.integers(1, 4))
    motions = [synth.make_motion(cfg, model, instance=i) for i in range(ni)]
    m = Model(ctx, model)
    clips = [Motion(m, mo) for mo in motions]
    orcs = [oracle.Restatement(model, mo) for mo in motions]
    first = [int(x) for x in rng.integers(0, cfg.n_frames, ni)]
    layout = capi.LAYOUT_SOA_POS_NRM if seed % 2 == 0 else capi.LAYOUT_INTERLEAVED_SOKOL32
    fr = Frames(m, ni, nf, layout)
    fr.update_range(clips, first, stride)
    ok = True

    def check(slot, orc, f):
        ref = orc.run_frame(f)
        good = same(fr.bone_matrices(slot), ref["skin"])
        if layout == capi.LAYOUT_SOA_POS_NRM:
            return good and same(fr.download(slot, capi.STREAM_POSITION), ref["pos"]) and same(fr.download(slot, capi.STREAM_NORMAL), ref["nrm"])
        return good and same(fr.download(slot, capi.STREAM_INTERLEAVED), orc.repack_sokol32())
    for i in range(ni):
        for k in range(nf):
            ok &= check(i * nf + k, orcs[i], first[i] + k * stride)
    per_slot = [int(x) for x in rng.integers(0, cfg.n_frames + 3, ni * nf)]
    fr.update(clips, per_slot)
    for i in range(ni):
        for k in range(nf):
            ok &= check(i * nf + k, orcs[i], per_slot[i * nf + k])
    fr.close()
    for o in orcs:
        o.close()
    return ok, f"crowd seed {seed}: {ni} instances x {nf} frames, stride {stride}"


def ext_case(seed):
    """extensions = 1 (parity unpinned): vertices that are not SDEF / QDEF stay bit-identical to the oracle; SDEF, QDEF
    and UV results are compared with the fp64 restatement of the documented formulas (tests/ext_reference.py, 1e-4)."""
    import ext_reference as xr
    rng = np.random.default_rng(17000 + seed)
    cfg = replace(synth.TINY_FULL, name=f"ext{seed}", config_id=1100 + seed, n_bones=int(rng.integers(12, 80)),
                  n_vertices=int(rng.integers(50, 1500)), n_vertex_morphs=int(rng.integers(0, 10)),
                  n_uv_morphs=int(rng.integers(0, 4)), n_group_morphs=int(rng.integers(0, 2)),
                  n_bone_morphs=int(rng.integers(0, 2)), ik_chains=int(rng.integers(0, 2)), n_frames=30,
                  binding=("coherent", "random")[int(rng.integers(0, 2))], stress=bool(rng.integers(0, 2)))
    if cfg.n_group_morphs and cfg.n_vertex_morphs + cfg.n_uv_morphs + cfg.n_bone_morphs < 3:
        cfg = replace(cfg, n_group_morphs=0)
    model = synth.add_material_morphs(synth.make_model(cfg), n_materials=int(rng.integers(1, 6)), seed=seed)
    motion = synth.make_motion(cfg, model)
    orc = oracle.Restatement(model, motion)
    m = Model(ctx, model, extensions=True)
    a = Motion(m, motion)
    frames = [int(x) for x in rng.integers(0, 33, 3)]
    fr = Frames(m, 1, len(frames))
    fr.update(a, frames)
    t_norm, ids, _ = orc.skinning()
    is_sdef = t_norm == capi.SKIN_SDEF
    is_qdef = model["skin_type"] == capi.SKIN_QDEF
    plain = ~(is_sdef | is_qdef)
    ok = True
    tol = dict(rtol=1e-4, atol=1e-4)
    for k, f in enumerate(frames):
        want = orc.run_frame(f)
        pos, nrm, skin = fr.download(k, capi.STREAM_POSITION), fr.download(k, capi.STREAM_NORMAL), fr.bone_matrices(k)
        ok &= same(skin, want["skin"]) and same(pos[plain], want["pos"][plain]) and same(nrm[plain], want["nrm"][plain])
        T = xr.bone_transforms(skin)
        rates = fr.morph_rates(k)
        dv, duv = xr.morph_images(model, rates)
        P = model["position"].astype(np.float64) + dv
        N = model["normal"].astype(np.float64)
        for i in np.flatnonzero(is_sdef)[:40]:
            p, n = xr.sdef(P[i], N[i], int(ids[i, 0]), int(ids[i, 1]), float(model["weight"][i, 0]), model["sdef_c"][i].astype(np.float64),
                           model["sdef_r0"][i].astype(np.float64), model["sdef_r1"][i].astype(np.float64), T)
            ok &= np.allclose(pos[i], p, **tol) and np.allclose(nrm[i], n, **tol)
        for i in np.flatnonzero(is_qdef)[:40]:
            p, n = xr.qdef(P[i], N[i], [int(x) for x in model["bone_id"][i]], model["weight"][i].astype(np.float64), T)
            ok &= np.allclose(pos[i], p, **tol) and np.allclose(nrm[i], n, **tol)
        ok &= np.allclose(fr.download(k, capi.STREAM_UV), model["uv"].astype(np.float64) + duv, rtol=1e-5, atol=1e-5)
        ok &= np.allclose(fr.material_images(k), xr.material_images(model, rates), rtol=1e-4, atol=1e-5)
    fr.close(); orc.close()
    return ok, str(cfg)


def main():
    global ctx
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    phase = sys.argv[3] if len(sys.argv) > 3 else "all"
    ctx = Context(0)
    bad = 0
    for name, fn in (("rig", rig_case), ("ik", ik_case), ("topo", topo_case), ("nest", nest_case), ("morph", morph_case), ("motion", motion_case), ("crowd", crowd_case), ("ext", ext_case)):
        if phase not in (name, "all"):
            continue
        n_bad = 0
        for seed in range(first, first + count):
            ok, what = fn(seed)
            if not ok:
                n_bad += 1
                print(f"MISMATCH {name} seed {seed}: {what}", flush=True)
        print(f"fuzz {name}: {count} configurations from seed {first}, {n_bad} mismatches", flush=True)
        bad += n_bad
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
