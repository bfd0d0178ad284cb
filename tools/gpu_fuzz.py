"""GPU box: randomized model / motion configurations beyond the seeds the test-suite pins, each compared bit-for-bit
with the CPU oracle (positions, normals, bone matrices, sampled poses, morph rates) in both output layouts.
usage: python tools/gpu_fuzz.py [first_seed] [count]"""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from dataclasses import replace
import numpy as np
import oracle
from simple_mmd_renderer_b200 import capi, synth
from simple_mmd_renderer_b200.poser import Context, Frames, Model, Motion

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ctx = Context(0)
bad = 0
for seed in range(first, first + count):
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
        ok &= np.array_equal(fr.download(k, capi.STREAM_POSITION).view(np.uint32), ref["pos"].view(np.uint32))
        ok &= np.array_equal(fr.download(k, capi.STREAM_NORMAL).view(np.uint32), ref["nrm"].view(np.uint32))
        ok &= np.array_equal(fr.bone_matrices(k).view(np.uint32), ref["skin"].view(np.uint32))
        ok &= np.array_equal(fr.bone_poses(k).view(np.uint32), ref["poses"].view(np.uint32))
        if ref["rates"].size:
            ok &= np.array_equal(fr.morph_rates(k).view(np.uint32), ref["rates"].view(np.uint32))
        ok &= np.array_equal(fi.download(k, capi.STREAM_INTERLEAVED).view(np.uint32), orc.repack_sokol32().view(np.uint32))
    if not ok:
        bad += 1
        print(f"MISMATCH seed {seed}: {cfg}", flush=True)
    fr.close(); fi.close(); orc.close()
print(f"fuzz: {count} configurations from seed {first}, {bad} mismatches", flush=True)
sys.exit(1 if bad else 0)
