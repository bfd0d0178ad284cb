"""GPU box: where the time of the interactive path (one slot, one frame per call) goes, for C1 (no IK) and C2 (two CCD
IK chains).  Host wall clock around synchronous sequences, 300 frames each, after a warm-up pass."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from simple_mmd_renderer_b200 import capi, synth
from simple_mmd_renderer_b200.poser import Context, Frames, Model, Motion

ctx = Context(0)
for cfg in (synth.C1, synth.C2):
    model = synth.make_model(cfg)
    motion = synth.make_motion(cfg, model)
    m = Model(ctx, model)
    mo = Motion(m, motion)
    fr = Frames(m, 1, 1)
    n = 300

    def stepwise(f):
        fr.reset_posing(); fr.seek_frame([mo], [f]); fr.pre_physics_posing(); fr.post_physics_posing(); fr.deform()

    def timed(fn, sync=True):
        for f in range(20):
            fn(f)
        ctx.synchronize()
        t0 = time.perf_counter()
        for f in range(n):
            fn(f)
            if sync:
                ctx.synchronize()
        ctx.synchronize()
        return (time.perf_counter() - t0) / n * 1e6

    def fused(f):
        fr.update([mo], [f])

    def fused_dl(f):
        fr.update([mo], [f]); fr.download(0, capi.STREAM_POSITION); fr.download(0, capi.STREAM_NORMAL)

    def stepwise_dl(f):
        stepwise(f); fr.download(0, capi.STREAM_POSITION); fr.download(0, capi.STREAM_NORMAL)

    print(f"{cfg.name}: stepwise+sync {timed(stepwise):.1f} us | fused+sync {timed(fused):.1f} us | "
          f"fused back-to-back {timed(fused, sync=False):.1f} us | stepwise+2 downloads {timed(stepwise_dl, sync=False):.1f} us | "
          f"fused+2 downloads {timed(fused_dl, sync=False):.1f} us", flush=True)
    ctx.set_profiling(True)
    for f in range(n):
        fused(f); ctx.synchronize()
    ms, launches = ctx.profile_read()
    print("   kernel us per frame (CUDA events): K1 %.1f  K2 %.1f  K3 %.1f" % tuple(1e3 * x / n for x in ms), flush=True)
    ctx.set_profiling(False)
