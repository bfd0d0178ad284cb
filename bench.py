#!/usr/bin/env python
"""bench.py — skinned vertex-frames/s of the libmmd deformation path on B200 (BASELINE.json metric).

One step = one pass of the whole hot path (VMD sampling K1 -> bone hierarchy / CCD IK / palette K2 -> morph
gather + skinning K3) over one batch of (instance, frame) slots of a synthetic PMX/VMD:

  workload C3 (default, headline): 1 M vertices, 1 k bones, 200 vertex morphs; `--frames-per-step` consecutive VMD
                         frames per step per GPU (the bake pattern of BASELINE configs[2]/[4]); N > 1 shards by frame
                         range, every rank fully independent ("scaling": "weak").
  workload C4:           512 instances of the 50 k-vertex model with independent clips, one frame each per
                         step, instances sharded across ranks ("scaling": "strong").
  workload C5:           offline bake: a 10 k-frame VMD on the 1 M-vertex model, frame range sharded across ranks, 64-frame
                         windows, every window's positions + normals gathered to rank 0 over NCCL ("scaling": "strong").
  workload C1 / C2:      one 50 k-vertex model (C2 adds SDEF/QDEF tags, UV/group/bone morphs, two CCD IK chains).

Whatever the headline workload, the same run also measures the other BASELINE configs briefly and reports them under
"also" (C4 strong scaling, the C5 bake with its gather, C3 with random bone binding, C1, C2 at several batch sizes) —
at every N, so that the driver's 1/2/4/8 scaling run carries them.

`--impl reference` times the reference's own CPU implementation (libmmd, compiled into oracle/_ref; else the
C restatement) on the host cores for the same workload.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "skinned_vertex_frames_per_sec"
UNIT = "vertex-frames/s"
BAKE_FRAMES = 10_000      # BASELINE configs[4]
BAKE_WINDOW = 64          # frames per fused update and per gather round


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mmdgpu", choices=["mmdgpu", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--frames-per-step", type=int, default=128, help="C1/C2/C3: VMD frames per step per GPU")
    ap.add_argument("--instances", type=int, default=512, help="C4: crowd size (whole job)")
    ap.add_argument("--layout", default="soa", choices=["soa", "sokol32"])
    ap.add_argument("--binding", default="coherent", choices=["coherent", "wide", "random"], help="vertex -> bone binding of the synthetic model")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short measurements of the other BASELINE configs")
    ap.add_argument("--also", default="C1,C2,C4,C5,C3_wide,C3_random", help="which secondary measurements to run")
    ap.add_argument("--bake-frames", type=int, default=BAKE_FRAMES)
    return ap.parse_args()


def algorithmic_bytes_per_vertex(model: dict, layout: str) -> float:
    """SURVEY 8(d), streaming model: 48 B static read + 24 B write + 4 B CSR row pointer + 16 B per morph entry
    (+8 B uv read +8 B wider record for the interleaved layout)."""
    e = float(model["n_vertex_morph_entries"]) / max(1, int(model["n_vertices"]))
    b = 48.0 + 24.0 + 4.0 + 16.0 * e
    if layout == "sokol32":
        b += 16.0
    return b


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_traffic_bytes(workload: str, layout: str, slots: int, binding: str = "coherent"):
    """dram read + write bytes per skin launch from the committed ncu capture of exactly this configuration, else None."""
    p = os.path.join(ROOT, "profiles", "skin_dram_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    key = f"{workload}/{layout}/{slots}" + ("" if binding == "coherent" else f"/{binding}")
    v = d.get(key)
    if v is None and layout == "soa" and slots == 128 and binding == "coherent":
        v = d.get(workload)   # round-1 key
    return v


def skin_roofline(model: dict, layout: str, slots: int, slots_per_cta: int, skin_ms: float, workload: str, binding: str) -> dict:
    """Roofline of the skinning kernel.

    `achieved` / `frac` are PHYSICAL: bytes that must cross the HBM interface per launch — every output record once
    (24 or 32 B per vertex-frame) plus the tile's static streams and morph table once per (tile, slot run) — divided by
    the measured launch time.  `traffic` is what ncu counted for the same launch (committed capture), and when it exists
    `frac` is computed from it instead of the model.  The SURVEY 8(d) streaming figure (static streams counted once per
    vertex-FRAME, which the kernel avoids by keeping a tile in registers across a slot run) is kept as *_streaming_model.
    """
    nv = int(model["n_vertices"])
    peak, peak_src = measured_peaks()
    e = float(model["n_vertex_morph_entries"]) / max(1, nv)
    out_b = 32.0 if layout == "sokol32" else 24.0
    static_b = 48.0 + 2.0 + 16.0 * e + (8.0 if layout == "sokol32" else 0.0)   # streams + PMX index + sliced-ELL entries (+uv)
    runs = max(1, (slots + max(1, slots_per_cta) - 1) // max(1, slots_per_cta))
    model_bytes = nv * (out_b * slots + static_b * runs)
    traffic = ncu_traffic_bytes(workload, layout, slots, binding)
    phys = float(traffic) if traffic else model_bytes
    b_alg = algorithmic_bytes_per_vertex(model, layout)
    t = skin_ms * 1e-3
    ach = phys / t / 1e9 if t > 0 else 0.0
    ach_stream = b_alg * nv * slots / t / 1e9 if t > 0 else 0.0
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
            "traffic": traffic, "kernel": "skin_kernel", "avg_launch_ms": skin_ms, "peak_source": peak_src,
            "bytes_per_launch": phys, "bytes_source": "ncu dram__bytes_read+write (profiles/skin_dram_traffic.json)" if traffic
            else "model: output once + static streams once per (tile, slot run)",
            "bytes_per_launch_model": model_bytes, "output_bytes_per_vertex_frame": out_b,
            "vertices_per_launch": nv * slots, "slots_per_cta_run": slots_per_cta,
            "frac_of_nominal_8TBs": ach / 8000.0,
            "algorithmic_bytes_per_vertex_streaming_model": b_alg,
            "achieved_streaming_model": ach_stream, "frac_streaming_model": ach_stream / peak if peak else None,
            "note": "frac is physical HBM traffic / time / measured copy peak; the streaming model counts static streams "
                    "once per vertex-frame (SURVEY 8d) although the kernel reads them once per 64-slot run, so it can exceed 1"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.proc = None
        self.path = None
        sel = str(device_index)
        try:
            import torch
            u = str(torch.cuda.get_device_properties(device_index).uuid)
            sel = u if u.startswith("GPU-") else "GPU-" + u
        except Exception:
            pass
        self.sel = sel

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.sel, f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
                except ValueError:
                    continue
                for nme, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no samples"}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def synth_config(workload: str, binding: str = "coherent"):
    from simple_mmd_renderer_b200 import synth
    cfg = synth.CONFIGS["C3" if workload == "C5" else workload]
    if workload == "C5":
        cfg = replace(cfg, name="C5")          # the C3 model; the 10 k-frame clip is built by bake_motion()
    if binding != "coherent":
        cfg = replace(cfg, binding=binding, name=f"{cfg.name}_{binding}")
    if os.environ.get("MMDGPU_BENCH_MORPHS"):      # experiments only: another number of vertex morphs (0 = no morph gather at all)
        n = int(os.environ["MMDGPU_BENCH_MORPHS"])
        cfg = replace(cfg, n_vertex_morphs=n, name=f"{cfg.name}_{n}morphs")
    return cfg


def bake_motion(cfg, model, n_frames: int):
    from simple_mmd_renderer_b200 import synth
    return synth.make_motion(replace(cfg, n_frames=n_frames), model)


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_session(model, motion):
    import oracle
    if not oracle.have_restatement() or (os.path.isdir("/root/reference") and not oracle.have_reference()):
        oracle.build()
    if oracle.have_reference():
        return oracle.Reference(model, motion), "reference"
    return oracle.Restatement(model, motion), "port"


def frames_for_step(step: int, n: int, clip_len: int) -> np.ndarray:
    first = (step * n) % max(1, clip_len)
    return ((first + np.arange(n)) % (clip_len + 1)).astype(np.uint32)


def slots_per_step(args, cfg) -> int:
    if args.workload == "C4":
        return args.instances // max(1, args.gpus)
    if args.workload == "C5":
        return BAKE_WINDOW
    return args.frames_per_step


def workload_config(args, cfg, model, n_slots):
    """Identical in both arms (the driver compares it): what is evaluated per step, not where."""
    e = float(model["n_vertex_morph_entries"]) / max(1, int(model["n_vertices"]))
    return {
        "workload": f"{cfg.name}: {cfg.n_vertices} vertices, {cfg.n_bones} bones, {int(model['n_morphs'])} morphs "
                    f"(e={e:.2f} morph entries/vertex), {cfg.n_frames}-frame VMD, physics off",
        "slots_per_step_per_gpu": int(n_slots), "layout": args.layout, "binding": cfg.binding,
        "l2": f"every step writes its slots' output ({n_slots * cfg.n_vertices * 24 / 1e6:.0f} MB per GPU; L2 is "
              "126 MB) between re-reads of the static streams; no separate flush",
    }


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from simple_mmd_renderer_b200 import synth
    cfg = synth_config(args.workload, args.binding)
    model = synth.make_model(cfg)
    motion = synth.make_motion(cfg, model)
    ses, kind = cpu_session(model, motion)
    T = host_threads()
    nv = int(model["n_vertices"])
    n_slots = slots_per_step(args, cfg)          # the same batch the GPU arm evaluates per step and per GPU
    times = []
    for s in range(args.warmup + args.steps):
        fr = frames_for_step(s, n_slots, cfg.n_frames)
        sec, _ = ses.time_frames(fr, T)
        if s >= args.warmup:
            times.append(sec)
    total = float(sum(times))
    value = n_slots * nv * args.steps / total
    sample = f"{n_slots} frames of {cfg.name} per step on {T} host threads (private Poser per thread, shared Model)"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload in ("C4", "C5") else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, cfg, model, n_slots),
        "where": "host: libmmd on the box's CPU threads",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": T, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------ own arm
class Env:
    """Process-wide state of one rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = f"cuda:{self.local}"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def timed_updates(env: Env, ctx, stream, step_fn, steps: int, warmup: int, profile: bool = True):
    """W untimed + K timed calls of step_fn(s) between barriers; returns (max-over-ranks ms, kernel ms[3], launches[3]).
    profile=False: no event pairs around the kernels (six event records per update are a visible share of a step of a
    few tens of microseconds); the kernel times then come from isolated_kernel_ms alone."""
    torch = env.torch
    for s in range(warmup):
        step_fn(s)
    ctx.synchronize()
    ctx.set_profiling(profile)
    ctx.profile_read()
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(steps):
        step_fn(warmup + s)
    e1.record(stream)
    env.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    kms, kn = ctx.profile_read()
    ctx.set_profiling(False)
    return ms, kms, kn


def isolated_kernel_ms(ctx, step_fn, base: int, reps: int = 3) -> dict:
    """Per-kernel time with nothing else running: every update is followed by a full synchronize, so the event pairs on
    the sampling / hierarchy streams are not stretched by the skinning kernel of the previous update."""
    ctx.synchronize()
    ctx.set_profiling(True)
    ctx.profile_read()
    for s in range(reps):
        step_fn(base + s)
        ctx.synchronize()
    kms, kn = ctx.profile_read()
    ctx.set_profiling(False)
    return {"pose_sample": kms[0] / reps, "hierarchy": kms[1] / reps, "skin": kms[2] / reps,
            "hierarchy_launches_per_update": kn[1] / reps,
            "note": "ms per update, each update run alone (synchronize between updates)"}


def measure_batch(env: Env, ctx, stream, workload: str, n_frames: int, steps: int = 10, warmup: int = 3, binding: str = "coherent",
                  model_cache: dict | None = None) -> dict:
    """Device-resident measurement of one model at one batch size on THIS rank (every rank does the same work)."""
    from simple_mmd_renderer_b200 import synth
    from simple_mmd_renderer_b200.poser import Frames, Model, Motion
    cfg = synth_config(workload, binding)
    key = (workload, binding)
    if model_cache is not None and key in model_cache:
        model, m, motion = model_cache[key]
    else:
        model = synth.make_model(cfg)
        m = Model(ctx, model)
        motion = Motion(m, synth.make_motion(cfg, model))
        if model_cache is not None:
            model_cache[key] = (model, m, motion)
    fr = Frames(m, 1, n_frames)

    def step(s):
        fr.update_range(motion, np.asarray([(s * 37) % max(1, cfg.n_frames - n_frames + 1)], np.uint32), 1)
    # no event pairs around the kernels inside the timed loop (their records are a visible share of steps this short); the
    # skinning kernel's own time comes from the isolated pass
    ms, _, _ = timed_updates(env, ctx, stream, step, steps, warmup, profile=False)
    nv = int(model["n_vertices"])
    iso = isolated_kernel_ms(ctx, step, warmup + steps, 2)
    skin_ms = iso["skin"]
    roof = skin_roofline(model, "soa", n_frames, fr.slots_per_cta, skin_ms, workload, binding)
    out = {"value": env.world * n_frames * nv * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "slots_per_step_per_gpu": n_frames, "scaling": "weak", "skin_ms_per_launch_isolated": skin_ms,
           "skin_frac_of_measured_hbm": roof["frac"], "skin_frac_streaming_model": roof["frac_streaming_model"],
           "kernel_ms_isolated": iso}
    fr.close()
    return out


def measure_crowd(env: Env, ctx, stream, total_inst: int, steps: int, warmup: int) -> dict:
    """BASELINE configs[3]: `total_inst` instances of the 50 k-vertex model with independent clips, contiguous instance
    blocks per rank, one frame per instance per step.  Strong scaling: the whole job is fixed."""
    from simple_mmd_renderer_b200 import shard, synth
    from simple_mmd_renderer_b200.poser import Frames, Model, Motion
    cfg = synth.CONFIGS["C4"]
    model = synth.make_model(cfg)
    m = Model(ctx, model)
    lo, hi = shard.split_range(total_inst, env.world, env.rank)
    motions = [Motion(m, synth.make_motion(cfg, model, instance=i)) for i in range(lo, hi)]
    fr = Frames(m, hi - lo, 1)
    rng = np.random.default_rng(99 + env.rank)
    clips = fr.anim_array(motions)            # marshalled once: the step is tens of microseconds of device work per rank at N = 8
    # the per-step inputs (one frame id per instance) are drawn before the timed loop: generating them is not the path
    frame_sets = [((s * 3 + rng.integers(0, cfg.n_frames, hi - lo)) % (cfg.n_frames + 1)).astype(np.uint32) for s in range(32)]

    def step(s):
        fr.update_range(clips, frame_sets[s % len(frame_sets)], 1)
    ms, _, _ = timed_updates(env, ctx, stream, step, steps, warmup, profile=False)
    nv = int(model["n_vertices"])
    out = {"value": total_inst * nv * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "instances": total_inst, "instances_per_gpu": hi - lo, "scaling": "strong", "n_gpus": env.world,
           "kernel_ms_isolated": isolated_kernel_ms(ctx, step, warmup + steps, 2),
           "batch": f"{total_inst} instances x 1 frame per step, independent clips, sharded by instance"}
    fr.close()
    for a in motions:
        a.close()
    m.close()
    return out


def measure_bake(env: Env, ctx, stream, m, model, cfg, n_frames_total: int, window: int) -> dict:
    """BASELINE configs[4]: a `n_frames_total`-frame clip baked on the 1 M-vertex model, contiguous frame ranges per rank,
    `window` frames per fused update.  Three passes over the whole clip:
      compute   every rank bakes its range, outputs stay on the device                       -> value_compute
      gather    + after every window, positions AND normals of all ranks go to rank 0 over NCCL (torch.distributed.gather
                on the library's output buffers, zero-copy; window k's gather overlaps window k+1's evaluation)
      p2p       + instead of NCCL, every rank's skinning kernel writes its window straight into rank 0's receive buffer
                (bulk stores over NVLink into peer memory bound with mmdgpu_frames_bind_output): compute and gather are
                one kernel; a 4-byte all-reduce per window orders it against the root
    `value` is the bake WITH the NCCL gather (whole clip, max over ranks)."""
    torch, dist = env.torch, env.dist
    from simple_mmd_renderer_b200 import capi, shard
    from simple_mmd_renderer_b200.poser import Frames, Motion
    nv = int(model["n_vertices"])
    motion = Motion(m, bake_motion(cfg, model, n_frames_total))
    lo, hi = shard.split_range(n_frames_total, env.world, env.rank)
    wins = list(shard.bake_windows(lo, hi, window))
    n_rounds = shard.n_windows(n_frames_total, env.world, window)
    frames = [Frames(m, 1, window), Frames(m, 1, window)]
    # outputs bound to plain torch tensors: contiguous [window, nv, 3], what NCCL sends without a staging copy
    outs = []
    for fr in frames:
        pos = torch.empty((window, nv, 3), dtype=torch.float32, device=env.dev)
        nrm = torch.empty((window, nv, 3), dtype=torch.float32, device=env.dev)
        fr.bind_output(capi.STREAM_POSITION, pos.data_ptr(), nv * 12)
        fr.bind_output(capi.STREAM_NORMAL, nrm.data_ptr(), nv * 12)
        outs.append((pos, nrm))
    recv = None
    if env.world > 1 and env.rank == 0:
        recv = [[(torch.empty((window, nv, 3), dtype=torch.float32, device=env.dev),
                  torch.empty((window, nv, 3), dtype=torch.float32, device=env.dev)) for _ in range(env.world)] for _ in range(2)]

    def evaluate(k):
        if k < len(wins):
            frames[k & 1].update_range(motion, np.asarray([wins[k][0]], np.uint32), 1)

    def run(mode):
        works = [None, None]
        with torch.cuda.stream(stream):
            env.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(n_rounds):
                b = k & 1
                if works[b] is not None:            # window k-2's gather must have read this buffer pair
                    for w in works[b]:
                        w.wait()                    # stream-level wait, the host does not block
                evaluate(k)
                if mode == "gather" and env.world > 1:
                    pos, nrm = outs[b]
                    works[b] = [dist.gather(pos, [r[0] for r in recv[b]] if env.rank == 0 else None, dst=0, async_op=True),
                                dist.gather(nrm, [r[1] for r in recv[b]] if env.rank == 0 else None, dst=0, async_op=True)]
            for ws in works:
                if ws is not None:
                    for w in ws:
                        w.wait()
            e1.record(stream)
            env.barrier()
        return env.max_over_ranks(e0.elapsed_time(e1))

    for k in range(min(2, len(wins))):               # warm-up: both frames objects, NCCL channels
        evaluate(k)
    ctx.synchronize()
    if env.world > 1:
        run("gather")
    ms_compute = run("compute")
    ms_gather = run("gather") if env.world > 1 else None
    total_vf = float(n_frames_total) * nv
    bytes_into_root = 0
    for r in range(1, env.world):
        rlo, rhi = shard.split_range(n_frames_total, env.world, r)
        bytes_into_root += (rhi - rlo) * nv * 24
    sent_padded = (env.world - 1) * n_rounds * window * nv * 24    # what NCCL actually moves (whole windows)
    seen = int(env.sum_over_ranks(1.0))
    out = {"value": total_vf / ((ms_gather if ms_gather else ms_compute) * 1e-3), "unit": UNIT, "scaling": "strong", "n_gpus": env.world,
           "frames": n_frames_total, "frames_per_gpu": hi - lo, "window": window, "windows_per_gpu": n_rounds,
           "bake_seconds_compute_only": ms_compute * 1e-3, "value_compute": total_vf / (ms_compute * 1e-3),
           "gather": None if ms_gather is None else {
               "bake_seconds_with_gather": ms_gather * 1e-3, "streams": "position + normal, every window",
               "bytes_into_root": int(bytes_into_root), "bytes_moved_incl_window_padding": int(sent_padded),
               "gbs_into_root": sent_padded / (ms_gather * 1e-3) / 1e9,
               "gbs_into_root_excess_over_compute": sent_padded / (max(ms_gather - ms_compute, 1e-6) * 1e-3) / 1e9,
               "comm_nranks_seen": seen, "collective": "torch.distributed.gather (NCCL), async, window k overlaps evaluation of k+1"},
           "comm_nranks_seen": seen,
           "note": "value = whole bake including the NCCL gather of every window to rank 0 (compute-only at N = 1: nothing to gather)"}
    # ---- fused variant: peers' skinning kernels store into rank 0's memory
    if env.world > 1:
        try:
            out["p2p_fused"] = bake_p2p(env, ctx, stream, frames, outs, motion, wins, n_rounds, window, nv, total_vf, sent_padded)
        except Exception as ex:   # IPC mapping is a platform capability; the NCCL numbers above stand without it
            out["p2p_fused"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
        env.barrier()
    for fr in frames:
        fr.close()
    motion.close()
    return out


def bake_p2p(env: Env, ctx, stream, frames, outs, motion, wins, n_rounds, window, nv, total_vf, sent_padded) -> dict:
    torch, dist = env.torch, env.dist
    from simple_mmd_renderer_b200 import capi, shard
    peer = shard.PeerWindows(ctx, env.world, env.rank, 2, window * nv * 3)   # [buffer][rank] x (pos, nrm) on rank 0
    flag = torch.zeros(1, dtype=torch.int32, device=env.dev)
    for b, fr in enumerate(frames):
        ppos, pnrm = peer.slot(b, env.rank)
        fr.bind_output(capi.STREAM_POSITION, ppos, nv * 12)
        fr.bind_output(capi.STREAM_NORMAL, pnrm, nv * 12)

    def run():
        with torch.cuda.stream(stream):
            env.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(n_rounds):
                if k < len(wins):
                    frames[k & 1].update_range(motion, np.asarray([wins[k][0]], np.uint32), 1)
                dist.all_reduce(flag)     # window k of every rank has landed on rank 0 / rank 0 may reuse buffer k & 1
            e1.record(stream)
            env.barrier()
        return env.max_over_ranks(e0.elapsed_time(e1))
    run()
    ms = run()
    # ---- check: every rank re-bakes its last window into local memory; rank 0 compares checksums of the bits
    for b, fr in enumerate(frames):
        fr.bind_output(capi.STREAM_POSITION, outs[b][0].data_ptr(), nv * 12)
        fr.bind_output(capi.STREAM_NORMAL, outs[b][1].data_ptr(), nv * 12)
    last = len(wins) - 1
    with torch.cuda.stream(stream):
        sums = torch.zeros(2, dtype=torch.int64, device=env.dev)
        if last >= 0:
            frames[last & 1].update_range(motion, np.asarray([wins[last][0]], np.uint32), 1)
            ctx.synchronize()
            sums[0] = outs[last & 1][0].view(torch.int32).sum(dtype=torch.int64)
            sums[1] = outs[last & 1][1].view(torch.int32).sum(dtype=torch.int64)
        info = torch.tensor([last], dtype=torch.int64, device=env.dev)
        all_sums = [torch.zeros_like(sums) for _ in range(env.world)]
        all_last = [torch.zeros_like(info) for _ in range(env.world)]
        dist.all_gather(all_sums, sums)
        dist.all_gather(all_last, info)
        ok = None
        if env.rank == 0:
            ok = True
            for r in range(env.world):
                lr = int(all_last[r].item())
                if lr < 0:
                    continue
                got0 = peer.tensor(lr & 1, r, 0, (window * nv * 3,)).view(torch.int32).sum(dtype=torch.int64)
                got1 = peer.tensor(lr & 1, r, 1, (window * nv * 3,)).view(torch.int32).sum(dtype=torch.int64)
                ok = ok and bool(got0 == all_sums[r][0]) and bool(got1 == all_sums[r][1])
        torch.cuda.synchronize()
    env.barrier()
    peer.close()
    return {"bake_seconds": ms * 1e-3, "value": total_vf / (ms * 1e-3), "gbs_into_root": sent_padded / (ms * 1e-3) / 1e9,
            "bit_identical_to_local_bake": ok,
            "how": "skin_kernel's cp.async.bulk shared->global stores target rank 0's buffer (CUDA IPC mapping, NVLink); "
                   "one 4-byte NCCL all-reduce per window orders producers and consumer"}


def run_mmdgpu(args):
    import torch
    import torch.distributed as dist
    from simple_mmd_renderer_b200 import capi, hostmem, lib, shard, synth
    from simple_mmd_renderer_b200.poser import Context, Frames, Model, Motion

    env = Env()
    world, rank, local = env.world, env.rank, env.local
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the deformation path has no CPU fallback")
    if not os.path.exists(lib.SO_PATH):
        if local == 0:
            lib.build_library()
    torch.cuda.set_device(local)
    numa = hostmem.bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()

    cfg = synth_config(args.workload, args.binding)
    model = synth.make_model(cfg)
    nv = int(model["n_vertices"])
    layout = capi.LAYOUT_SOA_POS_NRM if args.layout == "soa" else capi.LAYOUT_INTERLEAVED_SOKOL32
    stream = torch.cuda.Stream(device=local)
    ctx = Context(local, stream.cuda_stream)
    m = Model(ctx, model)
    out = None
    if args.workload == "C5":
        # the bake IS the step loop: one pass over the whole clip; `steps` / `warmup` are reported, not used
        bake = measure_bake(env, ctx, stream, m, model, cfg, args.bake_frames, BAKE_WINDOW)
        if rank == 0:
            out = {"metric": METRIC, "value": bake["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                   "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                   "data": "synthetic", "config": workload_config(args, cfg, model, BAKE_WINDOW), "bake": bake}
    else:
        out = headline(args, env, ctx, stream, m, model, cfg, layout, numa)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def headline(args, env: Env, ctx, stream, m, model, cfg, layout, numa):
    torch, dist = env.torch, env.dist
    from simple_mmd_renderer_b200 import capi, shard, synth
    from simple_mmd_renderer_b200.poser import Frames, Motion
    world, rank, local = env.world, env.rank, env.local
    nv = int(model["n_vertices"])
    if args.workload == "C4":
        lo, hi = shard.split_range(args.instances, world, rank)
        n_inst, n_frames = hi - lo, 1
        motions = [Motion(m, synth.make_motion(cfg, model, instance=i)) for i in range(lo, hi)]
    else:
        n_inst, n_frames = 1, args.frames_per_step
        motions = [Motion(m, synth.make_motion(cfg, model))]
    fr = Frames(m, n_inst, n_frames, layout)
    motion_objects = motions                  # keep the clips alive: the marshalled array only holds their handles
    motions = fr.anim_array(motion_objects)   # marshalled once
    slots = n_inst * n_frames
    rng = np.random.default_rng(1234 + rank)

    def first_frames(step):
        if args.workload == "C4":
            return ((step * 3 + rng.integers(0, cfg.n_frames, n_inst)) % (cfg.n_frames + 1)).astype(np.uint32)
        # bake pattern: rank r owns the r-th range of consecutive frames of this step
        base = (step * world + rank) * n_frames
        return np.asarray([base % max(1, cfg.n_frames - n_frames + 1)], np.uint32)

    def step(s):
        fr.update_range(motions, first_frames(s), 1)

    # ---- device-resident throughput (the headline `value`)
    sampler = ClockSampler(local) if rank == 0 else None
    for s in range(args.warmup):
        step(s)
    ctx.synchronize()
    ctx.set_profiling(True)
    ctx.profile_read()
    env.barrier()
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(args.steps):
        step(args.warmup + s)
    e1.record(stream)
    env.barrier()
    launches = ctx.launch_count - launches0
    total_ms = env.max_over_ranks(e0.elapsed_time(e1))
    kernel_ms, kernel_n = ctx.profile_read()
    ctx.set_profiling(False)
    total_slots = env.sum_over_ranks(float(slots))
    value = total_slots * nv * args.steps / (total_ms * 1e-3)
    iso = isolated_kernel_ms(ctx, step, args.warmup + args.steps)

    # ---- end to end through the C-ABI with host buffers: frame ids in, deformed vertex buffers out (pinned)
    e2e = None
    if not args.no_e2e:
        stream_ids = [capi.STREAM_POSITION, capi.STREAM_NORMAL] if args.layout == "soa" else [capi.STREAM_INTERLEAVED]
        per_stream = nv * (12 if args.layout == "soa" else 32) * slots
        host = [torch.empty(per_stream, dtype=torch.uint8, pin_memory=True) for _ in stream_ids]
        for h in host:
            h.zero_()
        h2d = 4 * n_inst
        d2h = per_stream * len(stream_ids)

        def e2e_step(s):
            fr.update_range(motions, first_frames(s), 1)          # H2D: frame ids (clips / model are resident)
            for sid, buf in zip(stream_ids, host):
                fr.download_async(0, slots, sid, buf.data_ptr(), per_stream)
            ctx.join_downloads()

        e2e_steps = max(3, min(args.steps, 10))
        for s in range(2):
            e2e_step(s)
        env.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for s in range(e2e_steps):
            e2e_step(2 + s)
        a1.record(stream)
        env.barrier()
        ems = env.max_over_ranks(a0.elapsed_time(a1))
        # the platform's ceiling for the same bytes: plain cudaMemcpyAsync of contiguous device memory into the same
        # pinned buffers, all ranks at once, no kernels (tools/d2h_ceiling.py is the standalone form)
        src = [torch.empty(per_stream, dtype=torch.uint8, device=env.dev) for _ in stream_ids]
        with torch.cuda.stream(stream):
            for h, d in zip(host, src):
                h.copy_(d, non_blocking=True)
            env.barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(3):
                for h, d in zip(host, src):
                    h.copy_(d, non_blocking=True)
            c1.record(stream)
            env.barrier()
        cms = env.max_over_ranks(c0.elapsed_time(c1)) / 3
        del src
        e2e_gbs = d2h * e2e_steps / (ems * 1e-3) / 1e9
        ceil_gbs = d2h / (cms * 1e-3) / 1e9
        e2e = {"value": total_slots * nv * e2e_steps / (ems * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "ms_per_step": ems / e2e_steps,
               "d2h_gbs_per_gpu": e2e_gbs, "d2h_ceiling_gbs_per_gpu": ceil_gbs, "frac_of_d2h_ceiling": e2e_gbs / ceil_gbs if ceil_gbs else None,
               "d2h_gbs_aggregate": e2e_gbs * world, "d2h_ceiling_gbs_aggregate": ceil_gbs * world,
               "note": "update_range + download of every slot's deformed buffer to pinned host memory, per GPU; the ceiling is "
                       "a bare cudaMemcpyAsync of the same bytes into the same buffers by all ranks at once",
               "numa": numa}
        del host

    clocks = sampler.stop() if sampler else None

    # ---- CPU baseline (rank 0, N = 1 only): libmmd itself on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        motion0 = synth.make_motion(cfg, model)
        ses, kind = cpu_session(model, motion0)
        T = host_threads()
        nfr = 8 * T if nv >= 500_000 else 64 * T
        sec, _ = ses.time_frames(frames_for_step(0, nfr, cfg.n_frames), T)
        cpu = {"value": nfr * nv / sec, "unit": UNIT, "cores": T, "kind": kind,
               "sample": f"{nfr} frames of {cfg.name} on {T} host threads, {sec:.2f} s wall"}
        # libmmd is single-threaded as shipped, and a release build would use -O3 / AVX2 / FMA
        n1 = max(4, nfr // T)
        sec1, _ = ses.time_frames(frames_for_step(0, n1, cfg.n_frames), 1)
        cpu["single_thread"] = {"value": n1 * nv / sec1, "cores": 1, "sample": f"{n1} frames, {sec1:.2f} s wall"}
        if kind == "reference":
            t0 = time.perf_counter()
            ses.repack_sokol32()
            cpu["repack_sokol32_ms_per_frame"] = 1e3 * (time.perf_counter() - t0)
            import oracle
            if oracle.have_reference_fast():
                fast = oracle.ReferenceFast(model, motion0)
                secf, _ = fast.time_frames(frames_for_step(0, nfr, cfg.n_frames), T)
                cpu["fast_build"] = {"value": nfr * nv / secf, "cores": T,
                                     "flags": "-O3 -march=x86-64-v3 (contraction on; not parity-grade)",
                                     "sample": f"{nfr} frames, {secf:.2f} s wall"}
                fast.close()
        ses.close()

    spc = fr.slots_per_cta
    fr.close()
    del motion_objects

    # ---- the other BASELINE configs, briefly, at every N (all ranks take part)
    also = None
    if not args.no_also:
        want = [w for w in args.also.split(",") if w]
        also = {}
        cache = {}

        def guarded(name, fn):
            try:
                also[name] = fn()
            except Exception as ex:          # a secondary measurement must not take the headline line down with it
                also[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                if world > 1:
                    raise                    # ...but ranks must not diverge inside collectives
        if "C5" in want and args.workload == "C3" and args.binding == "coherent":
            guarded("C5", lambda: measure_bake(env, ctx, stream, m, model, replace(cfg, name="C5"), args.bake_frames, BAKE_WINDOW))
        if "C4" in want and args.workload != "C4":
            guarded("C4", lambda: measure_crowd(env, ctx, stream, 512, 10, 3))
        if "C1" in want and args.workload != "C1":
            guarded("C1", lambda: measure_batch(env, ctx, stream, "C1", 512, model_cache=cache))
        if "C2" in want and args.workload != "C2":
            # two CCD IK chains per slot: small batches are bound by the latency of one solve chain, large ones by skinning
            for n in (128, 256, 512, 2048):
                guarded(f"C2_{n}", lambda n=n: measure_batch(env, ctx, stream, "C2", n, model_cache=cache))
        for _, mm, mo in cache.values():
            mo.close(); mm.close()
        cache.clear()
        if "C3_wide" in want and args.workload == "C3" and args.binding == "coherent":
            # bones drawn from +- 24 around the vertex's place in the skeleton: ~50 distinct bones per 512-vertex tile instead of
            # ~8, still staged in shared memory (real PMX vertex order is less bone-coherent than the headline model)
            guarded("C3_wide", lambda: measure_batch(env, ctx, stream, "C3", 128, steps=5, warmup=2, binding="wide"))
        if "C3_random" in want and args.workload == "C3" and args.binding == "coherent":
            # stress binding: every vertex picks its bones uniformly from all 1 k bones, so no tile-local palette fits and
            # the skinning kernel reads matrices from the slot's global palette (L2)
            guarded("C3_random", lambda: measure_batch(env, ctx, stream, "C3", 64, steps=5, warmup=2, binding="random"))

    if rank != 0:
        return None
    skin_ms = kernel_ms[2] / max(1, kernel_n[2])
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload == "C4" else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, cfg, model, slots),
        "where": "device-resident inputs and outputs (value); host buffers through the C-ABI (e2e)",
        "roofline": skin_roofline(model, args.layout, slots, spc, skin_ms, args.workload, args.binding),
        "kernel_ms": {"skin_per_launch_in_step": skin_ms, "isolated": iso,
                      "note": "sampling and hierarchy of update n+1 overlap the skinning kernel of update n on other streams; "
                              "their own durations are the isolated ones"},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if also is not None:
        out["also"] = also
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mmdgpu(args)


if __name__ == "__main__":
    main()
