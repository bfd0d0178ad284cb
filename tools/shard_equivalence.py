"""Multi-GPU shard equivalence on real GPUs: run under torchrun with N ranks.

Bake: every rank evaluates its contiguous frame range of one clip and the windows are gathered to rank 0 over NCCL
(shard.gather_window); rank 0 also bakes the whole range alone and compares bit-for-bit, and checks sampled frames
against the CPU oracle.  Crowd: every rank evaluates its block of instances; per-instance SHA-256 digests are
gathered and compared with rank 0's single-GPU run of all instances.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_equivalence.py
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simple_mmd_renderer_b200 import capi, shard, synth  # noqa: E402
from simple_mmd_renderer_b200.poser import Context, Frames, Model, Motion  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    cfg = synth.SMALL
    model = synth.make_model(cfg)
    motion = synth.make_motion(cfg, model)
    m = Model(ctx, model)
    a = Motion(m, motion)
    n_frames, window = 121, 16

    # ---- bake by frame range + NCCL gather of windows
    lo, hi = shard.split_range(n_frames, world, rank)
    chunks = list(shard.bake_windows(lo, hi, window))
    fr = Frames(m, 1, window)
    got = []
    for k in range(shard.n_windows(n_frames, world, window)):
        n_valid = 0
        if k < len(chunks):
            first, n_valid = chunks[k]
            fr.update_range(a, [first], 1)
            ctx.synchronize()
        local_t = shard.frames_as_tensor(fr, capi.STREAM_POSITION).contiguous()
        parts = shard.gather_window(local_t, n_valid, root=0)
        if rank == 0:
            got.append([(r, t.clone()) for r, t in parts])
    ok = True
    if rank == 0:
        sharded = torch.cat(shard.assemble(got, n_frames, world)).cpu().numpy()
        single = Frames(m, 1, n_frames)
        single.update_range(a, [0], 1)
        ctx.synchronize()
        alone = np.stack([single.download(k, capi.STREAM_POSITION) for k in range(n_frames)])
        same = bool((sharded.view(np.uint32) == alone.view(np.uint32)).all())
        import oracle
        orc = oracle.Restatement(model, motion)
        exact = all((orc.run_frame(f)["pos"].view(np.uint32) == sharded[f].view(np.uint32)).all() for f in (0, 59, 60, 120))
        print(f"bake: {world}-GPU frame-range shards gathered over NCCL == 1-GPU bake bit-for-bit: {same}; == CPU oracle on sampled frames: {exact}")
        ok = ok and same and exact

    # ---- bake with the FUSED gather: every rank's skinning kernel stores its windows straight into rank 0's memory
    #      (mmdgpu_frames_bind_output on a CUDA-IPC mapping of the root's buffer, NVLink), one small all-reduce per window
    try:
        nv = m.n_vertices
        n_rounds = shard.n_windows(n_frames, world, window)
        peer = shard.PeerWindows(ctx, world, rank, n_rounds, window * nv * 3)       # one buffer per round: nothing is reused
        flag = torch.zeros(1, dtype=torch.int32, device=f"cuda:{local}")
        stream = torch.cuda.current_stream()
        for k in range(n_rounds):
            if k < len(chunks):
                ppos, pnrm = peer.slot(k, rank)
                fr.bind_output(capi.STREAM_POSITION, ppos, nv * 12)
                fr.bind_output(capi.STREAM_NORMAL, pnrm, nv * 12)
                fr.update_range(a, [chunks[k][0]], 1)
            ctx.synchronize()
            dist.all_reduce(flag)
        fr.bind_output(capi.STREAM_POSITION, None)
        fr.bind_output(capi.STREAM_NORMAL, None)
        torch.cuda.synchronize()
        if rank == 0:
            same = True
            single_n = np.stack([single.download(k, capi.STREAM_NORMAL) for k in range(n_frames)])
            for r in range(world):
                rlo, rhi = shard.split_range(n_frames, world, r)
                for k, (first, n_valid) in enumerate(shard.bake_windows(rlo, rhi, window)):
                    gp = peer.tensor(k, r, 0, (window, nv, 3))[:n_valid].cpu().numpy()
                    gn = peer.tensor(k, r, 1, (window, nv, 3))[:n_valid].cpu().numpy()
                    same = same and bool((gp.view(np.uint32) == alone[first:first + n_valid].view(np.uint32)).all())
                    same = same and bool((gn.view(np.uint32) == single_n[first:first + n_valid].view(np.uint32)).all())
            print(f"bake, fused gather: {world} ranks' skinning kernels storing into rank 0's memory == 1-GPU bake bit-for-bit "
                  f"(positions and normals): {same}")
            ok = ok and same
        dist.barrier()
        peer.close()
    except Exception as ex:   # CUDA IPC is a platform capability
        if rank == 0:
            print(f"bake, fused gather: unavailable here ({type(ex).__name__}: {ex})")

    # ---- crowd by instance
    n_inst = 24
    clips = [synth.make_motion(cfg, model, instance=i) for i in range(n_inst)]
    first = (np.arange(n_inst, dtype=np.uint32) * 5) % 100
    ilo, ihi = shard.split_range(n_inst, world, rank)
    mine = Frames(m, ihi - ilo, 1)
    mine.update_range([Motion(m, clips[i]) for i in range(ilo, ihi)], first[ilo:ihi], 1)
    ctx.synchronize()
    digests = [hashlib.sha256(mine.download(i - ilo, capi.STREAM_POSITION).tobytes()).hexdigest() for i in range(ilo, ihi)]
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    if rank == 0:
        allinst = Frames(m, n_inst, 1)
        allinst.update_range([Motion(m, c) for c in clips], first, 1)
        ctx.synchronize()
        want = [hashlib.sha256(allinst.download(i, capi.STREAM_POSITION).tobytes()).hexdigest() for i in range(n_inst)]
        same = sum(gathered, []) == want
        print(f"crowd: {world}-GPU instance shards == 1-GPU run of all {n_inst} instances (SHA-256 per instance): {same}")
        ok = ok and same
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
