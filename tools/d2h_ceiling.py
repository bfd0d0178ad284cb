#!/usr/bin/env python
"""Platform ceiling of the device->host leg: N concurrent processes (one per GPU), each repeating ONE plain
`cudaMemcpyAsync` of a large device buffer into pinned host memory (no batch APIs, no kernels), all ranks
started together.  Prints one JSON line: per-GPU and aggregate GB/s, for the buffer sizes the e2e leg of bench.py
moves (3.07 GB per step per GPU on C3).

    python tools/d2h_ceiling.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/d2h_ceiling.py                         # 8 GPUs together

Variants timed: one 1-D copy of the whole buffer; the same bytes as `--pieces` 1-D copies (one per slot, what a
pitched 2-D copy degenerates to); a pitched cudaMemcpy2DAsync (what mmdgpu_frames_download_async used in round 1).
With --bind-numa the rank first binds itself (CPU affinity + preferred memory node) to the NUMA node `nvidia-smi topo`
reports for its GPU, so that the pinned buffer is node-local.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=3_072_000_000)
    ap.add_argument("--pieces", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bind-numa", dest="numa", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from simple_mmd_renderer_b200 import hostmem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bound = hostmem.bind_to_gpu_numa(local) if args.numa else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.bytes // args.pieces * args.pieces
    piece = n // args.pieces
    pitch = (piece + 6144 + 511) // 512 * 512          # device rows are padded like nv_pad slots
    dev = torch.empty(pitch * args.pieces, dtype=torch.uint8, device=f"cuda:{local}")
    dev.zero_()
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.zero_()                                        # touch: pages exist before timing
    st = torch.cuda.Stream(device=local)

    # Tensor.copy_(non_blocking=True) between a contiguous device tensor and a contiguous pinned tensor is exactly one
    # cudaMemcpyAsync(DeviceToHost) on the current stream
    def one_copy():
        with torch.cuda.stream(st):
            host.copy_(dev[:n], non_blocking=True)

    def piece_copies():
        with torch.cuda.stream(st):
            for i in range(args.pieces):
                host[i * piece:(i + 1) * piece].copy_(dev[i * pitch:i * pitch + piece], non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.reps):
            fn()
        e1.record(st)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return n * args.reps / (float(ms.item()) * 1e-3) / 1e9   # GB/s per GPU at the slowest rank

    res = {"one_1d_copy": timed(one_copy), f"{args.pieces}_1d_copies": timed(piece_copies)}
    if rank == 0:
        out = {"tool": "d2h_ceiling", "n_gpus": world, "bytes_per_copy": n, "reps": args.reps, "numa_bind": bound,
               "gbs_per_gpu": res, "gbs_aggregate": {k: v * world for k, v in res.items()},
               "host_cpus": len(os.sched_getaffinity(0)), "topology": hostmem.describe_topology()}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
