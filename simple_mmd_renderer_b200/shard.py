"""Multi-GPU decomposition of the deformation path: one process per GPU, no data-path collective.

The path shards where the reference's work is independent:
  * crowd  — instances are independent mmd::Poser objects: contiguous instance blocks per rank;
  * bake   — with physics off a frame is a pure function of its index (poser_impl.inl:362-377 resets all
             per-bone scratch every frame): contiguous frame ranges per rank.
The model (static streams, morph table, bone program) and the clips are replicated per GPU.  The only
communication is the optional final gather of baked vertex buffers, windowed because a whole bake does not fit
one GPU (10 k frames x 24 MB = 240 GB): `gather_window` moves one window of frames per rank to the root over
torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def split_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced block [lo, hi) of range(n) owned by `rank`; blocks differ by at most one item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world / rank")
    return n * rank // world, n * (rank + 1) // world


def owner_of(i: int, n: int, world: int) -> int:
    """Rank whose block contains item i (inverse of split_range)."""
    if not (0 <= i < n):
        raise ValueError("item out of range")
    r = (i * world) // n
    while split_range(n, world, r)[1] <= i:
        r += 1
    while split_range(n, world, r)[0] > i:
        r -= 1
    return r


def bake_windows(frame_lo: int, frame_hi: int, window: int):
    """(first_frame, n_frames) chunks a rank evaluates for its frame range; one chunk = one update_range call."""
    f = frame_lo
    while f < frame_hi:
        n = min(window, frame_hi - f)
        yield f, n
        f += n


def n_windows(n_frames_total: int, world: int, window: int) -> int:
    """Number of gather rounds: every rank must join every collective, so all use the longest rank's count."""
    longest = max(split_range(n_frames_total, world, r)[1] - split_range(n_frames_total, world, r)[0] for r in range(world))
    return (longest + window - 1) // window


class DeviceView:
    """Zero-copy torch view of a libmmdgpu device buffer (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, strides_bytes=None, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3,
                                         "strides": None if strides_bytes is None else tuple(int(x) for x in strides_bytes)}


def frames_as_tensor(frames, stream_id: int, n_slots: int | None = None):
    """torch tensor [n_slots, nv, 3] (SoA streams) or [n_slots, nv, 8] (interleaved) over the frames object's
    device output (slot stride as reported by mmdgpu_frames_device_ptr)."""
    import torch
    from . import capi
    ptr, stride = frames.device_ptr(stream_id)
    nv = frames.model.n_vertices
    width = 8 if stream_id == capi.STREAM_INTERLEAVED else 3
    n = frames.n_slots if n_slots is None else n_slots
    view = DeviceView(ptr, (n, nv, width), (stride, width * 4, 4))
    return torch.as_tensor(view, device=f"cuda:{frames.ctx.device}")


def gather_window(local, n_valid: int, root: int = 0, group=None):
    """Gather one window of baked frames to `root`.

    `local` is this rank's [window, nv, c] tensor (device tensor under NCCL, CPU tensor under gloo); only the first
    `n_valid` frames are meaningful (the last window of a rank may be short, and a rank that has run out of
    frames still joins with n_valid = 0).  Returns on the root a list of (rank, tensor[:n_valid_of_rank]) in rank
    order, elsewhere None.  Rank order == frame order because frame ranges are contiguous blocks by rank.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    counts[rank] = n_valid
    dist.all_reduce(counts, group=group)
    buf = local if local.is_contiguous() else local.contiguous()
    if rank == root:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=root, group=group)
        return [(r, parts[r][: int(counts[r].item())]) for r in range(world)]
    dist.gather(buf, None, dst=root, group=group)
    return None


def assemble(windows, n_frames_total: int, world: int):
    """Root-side: concatenate what `gather_window` returned over all windows into frame order (numpy / CPU torch)."""
    per_rank = [[] for _ in range(world)]
    for parts in windows:
        for r, t in parts:
            if t.shape[0]:
                per_rank[r].append(t)
    out = []
    for r in range(world):
        lo, hi = split_range(n_frames_total, world, r)
        got = sum(t.shape[0] for t in per_rank[r])
        if got != hi - lo:
            raise RuntimeError(f"rank {r} delivered {got} frames, expected {hi - lo}")
        out.extend(per_rank[r])
    return out


class BakeDriver:
    """Offline bake of one clip on one GPU (BASELINE configs[4], this rank's frame range): evaluates `window` frames
    per fused update and streams every window's deformed buffers to pinned host memory.  Two frames objects
    alternate: while the host sink consumes window k-1, the device evaluates window k and copies it out.  The driver
    waits per window (mmdgpu_frames_wait_downloads: that frames object's copies only), never for the whole context.

    sink(first_frame, n_frames, positions[n, nv, 3], normals[n, nv, 3]) is called with numpy views of the pinned
    buffers, valid until it returns.
    """

    def __init__(self, model, motion, window: int = 64):
        import torch
        from . import capi
        from .poser import Frames
        self.model, self.motion, self.window = model, motion, int(window)
        self.capi = capi
        self.frames = [Frames(model, 1, self.window), Frames(model, 1, self.window)]
        nbytes = self.window * model.n_vertices * 12
        self.host = [[torch.empty(max(nbytes, 16), dtype=torch.uint8, pin_memory=True) for _ in range(2)] for _ in range(2)]
        self.nbytes = nbytes
        self.issued = 0   # windows whose evaluation + copy have been queued (tests look at this from inside the sink)

    def run(self, frame_lo: int, frame_hi: int, sink):
        nv = self.model.n_vertices
        pending = None  # (buffer index, first frame, n)
        self.issued = 0
        for first, n in bake_windows(frame_lo, frame_hi, self.window):
            b = self.issued & 1
            fr = self.frames[b]
            # slots past n hold frames beyond the range: evaluated, not copied.  The skinning kernel of this update waits
            # on the device for this object's previous copy (window k-2), which the sink has long consumed.
            fr.update_range(self.motion, [first], 1)
            if n and self.nbytes:
                for sid, buf in zip((self.capi.STREAM_POSITION, self.capi.STREAM_NORMAL), self.host[b]):
                    fr.download_async(0, n, sid, buf.data_ptr(), n * nv * 12)
            self.issued += 1
            if pending is not None:
                self._deliver(pending, sink)       # window k-1 on the host while window k runs on the device
            pending = (b, first, n)
        if pending is not None:
            self._deliver(pending, sink)

    def _deliver(self, pending, sink):
        import torch
        b, first, n = pending
        self.frames[b].wait_downloads()            # this window's two copies, nothing else
        nv = self.model.n_vertices
        pos = self.host[b][0][: n * nv * 12].view(dtype=torch.float32).reshape(n, nv, 3).numpy()
        nrm = self.host[b][1][: n * nv * 12].view(dtype=torch.float32).reshape(n, nv, 3).numpy()
        sink(first, n, pos, nrm)

    def close(self):
        for fr in self.frames:
            fr.close()


class PeerWindows:
    """Receive buffer of the fused bake gather: `n_buffers` x `world` windows of (position, normal) planes in rank 0's
    memory, mapped into every rank (CUDA IPC over NVLink) so that each rank's skinning kernel writes its window straight
    into its place (mmdgpu_frames_bind_output).  Layout: [buffer][rank][plane][floats_per_plane] float32.
    The 64-byte handle travels over torch.distributed; ordering of producers and the consumer is the caller's job."""

    def __init__(self, ctx, world: int, rank: int, n_buffers: int, floats_per_plane: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from .lib import check
        self.ctx, self.world, self.rank = ctx, world, rank
        self.n_buffers, self.fpp = int(n_buffers), int(floats_per_plane)
        self.total_bytes = self.n_buffers * world * 2 * self.fpp * 4
        p = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self.opened = rank != 0
        if rank == 0:
            check(ctx.lib.mmdgpu_peer_buffer_create(ctx.h, self.total_bytes, C.byref(p), handle), ctx.h)
        t = torch.tensor(list(handle), dtype=torch.uint8, device=f"cuda:{ctx.device}")
        dist.broadcast(t, src=0, group=group)
        if rank != 0:
            handle = (C.c_ubyte * 64)(*t.cpu().tolist())
            check(ctx.lib.mmdgpu_peer_buffer_open(ctx.h, handle, C.byref(p)), ctx.h)
        self.base = int(p.value)

    def slot(self, buffer: int, rank: int) -> tuple[int, int]:
        """Device addresses (position plane, normal plane) of `rank`'s window in `buffer`."""
        off = ((buffer * self.world + rank) * 2) * self.fpp * 4
        return self.base + off, self.base + off + self.fpp * 4

    def tensor(self, buffer: int, rank: int, plane: int, shape):
        """torch view of one plane (any rank can look, but only rank 0 reads local memory)."""
        import torch
        addr = self.slot(buffer, rank)[plane]
        return torch.as_tensor(DeviceView(addr, shape), device=f"cuda:{self.ctx.device}")

    def close(self):
        from .lib import check
        import ctypes as C
        if getattr(self, "base", 0):
            check(self.ctx.lib.mmdgpu_peer_buffer_release(self.ctx.h, C.c_void_p(self.base), 1 if self.opened else 0), self.ctx.h)
            self.base = 0
