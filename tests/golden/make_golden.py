"""Generate tests/golden/*.npz from the reference itself (libmmd compiled into oracle/_ref/libmmd_ref.so).

Run in the build container, where /root/reference is mounted:   python tests/golden/make_golden.py
The fixtures pin the C restatement (oracle/mmd_oracle.c) and the CUDA path on machines where the reference
cannot be built (the GPU box).  Per (config, frame): SHA-256 of the full position / normal / skinning /
local-matrix / pose / rate arrays, the complete bone matrices, and a strided vertex sample.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from simple_mmd_renderer_b200 import synth  # noqa: E402
import oracle  # noqa: E402

CASES = {
    "tiny": [0, 1, 13, 30, 59, 60, 200],
    "tiny_full": [0, 1, 2, 7, 17, 30, 42, 45, 63, 88, 89, 90, 500],
    "small": [0, 59, 119],
    "ik_zoo": list(range(0, 41, 3)),     # synth.make_ik_zoo(): every CCD IK branch
    "ik_nested": list(range(0, 41, 4)),  # synth.make_ik_nested(): solves inside solves (poser_impl.inl:203-206, :303)
}
STRIDE = 53
TIMES = [0.0, 1.0 / 30.0, 0.0123, 0.5, 0.7777, 1.99, 2.5, 17.0]   # seconds, MotionPlayer::SeekTime


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def input_digest(model: dict, motion: dict) -> str:
    h = hashlib.sha256()
    for d in (model, motion):
        for k in sorted(d):
            v = d[k]
            if v is None:
                continue
            h.update(k.encode())
            h.update(np.ascontiguousarray(v).tobytes() if isinstance(v, np.ndarray) else str(int(v)).encode())
    return h.hexdigest()


def main():
    oracle.build()
    assert oracle.have_reference(), "the reference harness needs /root/reference"
    for name, frames in CASES.items():
        if name == "ik_zoo":
            model, motion = synth.make_ik_zoo()
        elif name == "ik_nested":
            model, motion = synth.make_ik_nested()
        else:
            cfg = synth.CONFIGS[name]
            model = synth.make_model(cfg)
            motion = synth.make_motion(cfg, model)
        ref = oracle.Reference(model, motion)
        out = {"frames": np.asarray(frames, np.uint32), "input_digest": np.asarray(input_digest(model, motion)),
               "stride": np.asarray(STRIDE)}
        t, ids, w = ref.skinning()
        out["norm_type"], out["norm_ids"], out["norm_w"] = t, ids, w
        fix, order = ref.ik_class()
        out["ik_fix"], out["ik_order"] = fix, order
        for f in frames:
            r = ref.run_frame(f)
            for k in ("pos", "nrm", "skin", "local", "poses", "rates"):
                out[f"f{f}_{k}_sha"] = np.asarray(sha(r[k]))
            out[f"f{f}_skin"] = r["skin"]
            out[f"f{f}_poses"] = r["poses"]
            out[f"f{f}_rates"] = r["rates"]
            out[f"f{f}_pos_s"] = r["pos"][::STRIDE].copy()
            out[f"f{f}_nrm_s"] = r["nrm"][::STRIDE].copy()
            out[f"f{f}_sokol32_sha"] = np.asarray(sha(ref.repack_sokol32()))
        out["times"] = np.asarray(TIMES, np.float64)
        for i, t in enumerate(TIMES):
            r = ref.run_time(t)
            for k in ("pos", "nrm", "skin", "poses", "rates"):
                out[f"t{i}_{k}_sha"] = np.asarray(sha(r[k]))
            out[f"t{i}_poses"] = r["poses"]
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
