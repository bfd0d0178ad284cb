#!/usr/bin/env python
"""Offline bake of a VMD clip on a PMX model (the BASELINE configs[4] pattern): every frame of the clip through the
fused update, deformed buffers streamed to the host window by window, optionally sharded over ranks by frame range
(torchrun).  Usage:

    python examples/bake_clip.py model.pmx motion.vmd out_prefix [--window 64]
    python -m torch.distributed.run --nproc-per-node 8 examples/bake_clip.py model.pmx motion.vmd out_prefix

Each rank writes <out_prefix>.rank<r>.npz with `first_frame`, `positions[n, nv, 3]`, `normals[n, nv, 3]` of its
contiguous frame range (libmmd's pose_image.coordinates / .normals per frame, main.cpp:827-844)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_mmd_renderer_b200 import shard  # noqa: E402
from simple_mmd_renderer_b200.poser import Context, Model, Motion  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pmx")
    ap.add_argument("vmd")
    ap.add_argument("out_prefix")
    ap.add_argument("--window", type=int, default=64)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    model = Model(ctx, pmx_bytes=open(args.pmx, "rb").read())
    motion = Motion(model, vmd_bytes=open(args.vmd, "rb").read())
    lo, hi = shard.split_range(motion.GetLength() + 1, world, rank)   # frames are pure functions of their index: no exchange
    pos = np.empty((hi - lo, model.n_vertices, 3), np.float32)
    nrm = np.empty_like(pos)

    def sink(first, n, p, q):
        pos[first - lo:first - lo + n] = p
        nrm[first - lo:first - lo + n] = q

    shard.BakeDriver(model, motion, window=args.window).run(lo, hi, sink)
    np.savez(f"{args.out_prefix}.rank{rank}.npz", first_frame=lo, positions=pos, normals=nrm)
    print(f"rank {rank}: frames [{lo}, {hi}) of {model.n_vertices} vertices baked")


if __name__ == "__main__":
    main()
