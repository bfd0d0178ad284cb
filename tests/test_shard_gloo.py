"""Multi-GPU host logic on CPU: sharding arithmetic, and the windowed gather of baked frames over a world_size-2
gloo group (the NCCL path uses the same code with device tensors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simple_mmd_renderer_b200 import shard


@pytest.mark.parametrize("n,world", [(512, 1), (512, 2), (512, 8), (10_000, 8), (7, 8), (10, 3), (0, 4)])
def test_split_range_is_a_partition(n, world):
    blocks = [shard.split_range(n, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for (a, b), (c, d) in zip(blocks, blocks[1:]):
        assert b == c and a <= b and c <= d
    sizes = [b - a for a, b in blocks]
    assert max(sizes) - min(sizes) <= 1
    for i in range(0, n, max(1, n // 50)):
        r = shard.owner_of(i, n, world)
        assert blocks[r][0] <= i < blocks[r][1]


def test_bake_windows_cover_the_range_once():
    for lo, hi, w in [(0, 10_000, 64), (1250, 2500, 64), (5, 6, 64), (3, 3, 8)]:
        seen = []
        for f, n in shard.bake_windows(lo, hi, w):
            assert 1 <= n <= w
            seen.extend(range(f, f + n))
        assert seen == list(range(lo, hi))
    assert shard.n_windows(10_000, 8, 64) == 20
    assert shard.n_windows(7, 8, 64) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frame_payload(f: int, nv: int) -> torch.Tensor:
    """Stand-in for a baked frame: value encodes (frame, vertex, component)."""
    v = torch.arange(nv, dtype=torch.float32).reshape(nv, 1)
    c = torch.arange(3, dtype=torch.float32).reshape(1, 3)
    return f * 1000.0 + v + c / 10.0


def _worker(rank, world, port, n_frames, window, nv, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.split_range(n_frames, world, rank)
        chunks = list(shard.bake_windows(lo, hi, window))
        rounds = shard.n_windows(n_frames, world, window)
        got = []
        for k in range(rounds):
            local = torch.zeros(window, nv, 3)
            n_valid = 0
            if k < len(chunks):
                f0, n_valid = chunks[k]
                for i in range(n_valid):
                    local[i] = _frame_payload(f0 + i, nv)
            parts = shard.gather_window(local, n_valid, root=0)
            if rank == 0:
                got.append([(r, t.clone()) for r, t in parts])
            else:
                assert parts is None
        if rank == 0:
            frames = torch.cat(shard.assemble(got, n_frames, world))
            torch.save(frames, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,window", [(37, 8), (5, 4), (16, 16)])
def test_windowed_gather_world2_gloo(tmp_path, n_frames, window):
    world, nv = 2, 11
    out = str(tmp_path / "frames.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_frames, window, nv, out), nprocs=world, join=True)
    frames = torch.load(out)
    assert frames.shape == (n_frames, nv, 3)
    want = torch.stack([_frame_payload(f, nv) for f in range(n_frames)])
    assert torch.equal(frames, want), "gathered bake is not in frame order"


def _crowd_worker(rank, world, port, n_inst, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.split_range(n_inst, world, rank)
        mine = torch.arange(lo, hi, dtype=torch.int64)
        total = torch.tensor([mine.numel()], dtype=torch.int64)
        dist.all_reduce(total)
        assert int(total) == n_inst
        np.save(os.path.join(out_dir, f"r{rank}.npy"), mine.numpy())
    finally:
        dist.destroy_process_group()


def test_crowd_instances_are_sharded_without_overlap_world2_gloo(tmp_path):
    mp.spawn(_crowd_worker, args=(2, _free_port(), 513, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(np.concatenate([a, b]), np.arange(513))
