"""GPU tests of the round-2 entry points: clip identity across destroy / create, ordering of skinning against
asynchronous downloads, caller-owned output buffers (the GL-free half of the zero-copy hand-off),
per-window waits of the bake driver.  All comparisons are bit-exact against the oracle (libmmd where its harness is built)."""
import time
from dataclasses import replace

import numpy as np
import pytest

from conftest import assert_bitwise, synth_case

pytestmark = pytest.mark.gpu

from simple_mmd_renderer_b200 import capi, synth  # noqa: E402
from simple_mmd_renderer_b200.poser import Frames, MmdGpuError, Model, Motion  # noqa: E402


def _oracle(model, motion):
    import oracle
    if oracle.have_reference():
        return oracle.Reference(model, motion)
    return oracle.Restatement(model, motion)


def test_clip_destroyed_and_another_created_is_not_mistaken_for_the_first(ctx):
    """destroy(A); create(B) usually hands B the heap block A had: the frames object must still upload B's arrays."""
    cfg, model, motion_a = synth_case("tiny_full")
    m = Model(ctx, model)
    fr = Frames(m, 1, 3)
    frames = [5, 31, 77]
    for round_ in range(4):
        motion = synth.make_motion(cfg, model, instance=round_)     # a different clip every round
        a = Motion(m, motion)
        for _ in range(10):                                         # every rotating state copy sees this clip
            fr.update(a, frames)
        orc = _oracle(model, motion)
        for k, f in enumerate(frames):
            ref = orc.run_frame(f)
            assert_bitwise(fr.bone_poses(k), ref["poses"], f"round {round_} frame {f} poses")
            assert_bitwise(fr.download(k, capi.STREAM_POSITION), ref["pos"], f"round {round_} frame {f} positions")
        a.close()                                                   # mmdgpu_animation_destroy: the next clip reuses the block


def test_rejected_update_leaves_the_last_good_result_selected(ctx):
    cfg, model, motion = synth_case("tiny")
    m = Model(ctx, model)
    a = Motion(m, motion)
    fr = Frames(m, 1, 1)
    fr.update(a, [12])
    want = fr.download(0, capi.STREAM_POSITION)
    skin = fr.bone_matrices(0)
    other = Model(ctx, model)
    foreign = Motion(other, motion)                                 # bound to a different model handle
    with pytest.raises(MmdGpuError):
        fr.update(foreign, [40])
    assert_bitwise(fr.bone_matrices(0), skin, "skinning matrices after a rejected update")
    assert_bitwise(fr.download(0, capi.STREAM_POSITION), want, "positions after a rejected update")
    fr.update(a, [12])
    assert_bitwise(fr.download(0, capi.STREAM_POSITION), want, "positions after the next good update")


def test_update_right_after_download_async_does_not_tear_the_copy(ctx):
    """One frames object: download of update n is still in flight when update n+1 is issued.  The skinning kernel of
    n+1 must wait for the copy; what arrives on the host is update n, bit for bit."""
    import torch
    cfg, model, motion = synth_case("C1")
    m = Model(ctx, model)
    a = Motion(m, motion)
    n = 96
    fr = Frames(m, 1, n)
    nv = m.n_vertices
    host = [torch.empty(n * nv * 12, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    orc = _oracle(model, motion)
    for first in (0, 100, 200):
        fr.update_range(a, [first], 1)
        for sid, buf in zip((capi.STREAM_POSITION, capi.STREAM_NORMAL), host):
            fr.download_async(0, n, sid, buf.data_ptr(), n * nv * 12)
        fr.update_range(a, [(first + 150) % 300], 1)       # overwrites the same output planes
        fr.update_range(a, [(first + 37) % 300], 1)
        fr.wait_downloads()
        assert fr.downloads_done()
        pos = host[0].view(torch.float32).reshape(n, nv, 3).numpy()
        nrm = host[1].view(torch.float32).reshape(n, nv, 3).numpy()
        for k in (0, 1, n // 2, n - 1):
            ref = orc.run_frame(first + k)
            assert_bitwise(pos[k], ref["pos"], f"window at {first}, slot {k} positions")
            assert_bitwise(nrm[k], ref["nrm"], f"window at {first}, slot {k} normals")
    ctx.synchronize()


@pytest.mark.parametrize("nv", [3000, 3001, 3002, 3003, 517])
def test_bound_interleaved_output_is_main_cpp_repack(ctx, nv):
    """mmdgpu_frames_bind_output(INTERLEAVED): the skinning kernel writes main.cpp:50-54 records straight into a
    caller-owned device buffer of exactly nv records per slot (what a mapped GL vertex buffer is), nothing beyond it.
    Compared bit-for-bit with libmmd's Deform + the repack of main.cpp:838-859."""
    import torch
    cfg = replace(synth.TINY_FULL, n_vertices=nv, name=f"tiny_full_{nv}")
    model = synth.make_model(cfg)
    motion = synth.make_motion(cfg, model)
    m = Model(ctx, model)
    a = Motion(m, motion)
    n_slots = 3
    fr = Frames(m, 1, n_slots, capi.LAYOUT_INTERLEAVED_SOKOL32)
    stride = nv * 32
    guard = 4096
    buf = torch.full((n_slots * stride + guard,), 0xAB, dtype=torch.uint8, device="cuda:0")
    fr.bind_output(capi.STREAM_INTERLEAVED, buf.data_ptr(), stride)
    frames = [4, 40, 88]
    fr.update(a, frames)
    ctx.synchronize()
    got = buf.cpu().numpy()
    assert (got[n_slots * stride:] == 0xAB).all(), "the kernel wrote past the bound buffer"
    orc = _oracle(model, motion)
    for k, f in enumerate(frames):
        orc.run_frame(f)
        want = orc.repack_sokol32()
        rec = got[k * stride:(k + 1) * stride].view(np.float32).reshape(nv, 8)
        assert_bitwise(rec, want, f"nv {nv} frame {f} sokol32 records in the bound buffer")
    # the download entry points follow the binding
    assert_bitwise(fr.download(1, capi.STREAM_INTERLEAVED), got[stride:2 * stride].view(np.float32).reshape(nv, 8), "download follows binding")
    # unbind: the library's own buffer again
    fr.bind_output(capi.STREAM_INTERLEAVED, None)
    buf.fill_(0xCD)
    fr.update(a, frames)
    ctx.synchronize()
    assert (buf.cpu().numpy() == 0xCD).all(), "an unbound buffer must not be written"
    orc.run_frame(frames[2])
    assert_bitwise(fr.download(2, capi.STREAM_INTERLEAVED), orc.repack_sokol32(), "library buffer after unbinding")


@pytest.mark.parametrize("nv", [3000, 3001, 3002, 3003])
def test_bound_soa_outputs_with_ragged_last_tile(ctx, nv):
    """POSITION / NORMAL planes bound to caller memory, nv records per slot: the last tile's bulk copy stops at the
    16-byte boundary below nv * 12 and the remaining floats are stored individually."""
    import torch
    cfg = replace(synth.TINY_FULL, n_vertices=nv, name=f"tiny_full_{nv}")
    model = synth.make_model(cfg)
    motion = synth.make_motion(cfg, model)
    m = Model(ctx, model)
    a = Motion(m, motion)
    n_slots = 5
    fr = Frames(m, 1, n_slots)
    stride = (nv * 12 + 15) // 16 * 16
    guard = 1024
    pos = torch.full((n_slots * stride + guard,), 0xAB, dtype=torch.uint8, device="cuda:0")
    nrm = torch.full((n_slots * stride + guard,), 0xAB, dtype=torch.uint8, device="cuda:0")
    fr.bind_output(capi.STREAM_POSITION, pos.data_ptr(), stride)
    fr.bind_output(capi.STREAM_NORMAL, nrm.data_ptr(), stride)
    frames = [0, 9, 33, 61, 89]
    fr.update(a, frames)
    ctx.synchronize()
    gp, gn = pos.cpu().numpy(), nrm.cpu().numpy()
    orc = _oracle(model, motion)
    for k, f in enumerate(frames):
        ref = orc.run_frame(f)
        assert_bitwise(gp[k * stride:k * stride + nv * 12].view(np.float32).reshape(nv, 3), ref["pos"], f"nv {nv} frame {f} positions")
        assert_bitwise(gn[k * stride:k * stride + nv * 12].view(np.float32).reshape(nv, 3), ref["nrm"], f"nv {nv} frame {f} normals")
        assert (gp[k * stride + nv * 12:(k + 1) * stride] == 0xAB).all(), "bytes between slots were written"
    assert (gp[n_slots * stride:] == 0xAB).all() and (gn[n_slots * stride:] == 0xAB).all()
    with pytest.raises(MmdGpuError):
        fr.bind_output(capi.STREAM_POSITION, pos.data_ptr() + 4, stride)      # misaligned
    with pytest.raises(MmdGpuError):
        fr.bind_output(capi.STREAM_INTERLEAVED, pos.data_ptr(), stride)       # wrong layout


@pytest.mark.parametrize("nv", [3000, 3002])
def test_outputs_bound_to_page_locked_host_memory(ctx, nv):
    """POSITION / NORMAL bound to pinned HOST memory: the skinning kernel's stores cross PCIe themselves, the host reads
    the planes after mmdgpu_frames_wait_skinning - no device-to-host copy (what mmdgpu::Poser::pose_image does).  Both
    the fused update and the step-wise Deform are covered; bytes outside the records stay untouched."""
    import torch
    cfg = replace(synth.TINY_FULL, n_vertices=nv, name=f"tiny_full_{nv}")
    model = synth.make_model(cfg)
    motion = synth.make_motion(cfg, model)
    m = Model(ctx, model)
    a = Motion(m, motion)
    orc = _oracle(model, motion)
    stride = (nv * 12 + 15) // 16 * 16
    for n_slots in (1, 3):
        fr = Frames(m, 1, n_slots)
        guard = 256
        host = torch.full((2 * n_slots * stride + guard,), 0xAB, dtype=torch.uint8).pin_memory()
        fr.bind_output(capi.STREAM_POSITION, host.data_ptr(), stride)
        fr.bind_output(capi.STREAM_NORMAL, host.data_ptr() + n_slots * stride, stride)
        frames = [7, 40, 88][:n_slots]
        for rep in range(2):
            if rep == 0:
                fr.update(a, frames)
            else:   # libmmd's call sequence, one call at a time
                frames = [f + 3 for f in frames]
                fr.reset_posing(); fr.seek_frame(a, frames); fr.pre_physics_posing(); fr.post_physics_posing(); fr.deform()
            fr.wait_skinning()
            g = host.numpy()
            for k, f in enumerate(frames):
                ref = orc.run_frame(f)
                assert_bitwise(g[k * stride:k * stride + nv * 12].view(np.float32).reshape(nv, 3), ref["pos"], f"host-bound positions, frame {f}")
                o = (n_slots + k) * stride
                assert_bitwise(g[o:o + nv * 12].view(np.float32).reshape(nv, 3), ref["nrm"], f"host-bound normals, frame {f}")
                assert (g[k * stride + nv * 12:(k + 1) * stride] == 0xAB).all()
            assert (g[2 * n_slots * stride:] == 0xAB).all()
        fr.close()
    with pytest.raises(MmdGpuError):
        Frames(m, 1, 1).bind_output(capi.STREAM_POSITION, np.zeros(nv * 3 + 8, np.float32).ctypes.data // 16 * 16 + 16, stride)   # pageable memory


def test_gl_interop_of_the_bound_output_needs_a_gl_stack():
    """SURVEY 8f-1, GL half: cudaGraphicsGLRegisterBuffer on sokol's vertex buffer + mmdgpu_frames_bind_output
    ("Zero-copy hand-off to the GL vertex buffer").  It needs libEGL (surfaceless context) and libGL on the
    GPU box; the boxes of this project have neither, so only the GL-free half is exercised
    (test_bound_interleaved_output_is_main_cpp_repack: a plain device allocation stands in for the mapped GL buffer)."""
    import ctypes.util
    missing = [n for n in ("EGL", "GL") if ctypes.util.find_library(n) is None]
    if missing:
        pytest.skip("no " + " / ".join("lib" + n for n in missing) + " on this box: a surfaceless EGL context cannot be created, "
                    "so cudaGraphicsGLRegisterBuffer has no buffer object to register")
    pytest.skip("libEGL and libGL are present, but this repository carries no EGL / GL loader: run the snippet of INTEGRATION.md "
                "against the renderer's own sokol context")


def test_bake_driver_hands_a_window_to_the_sink_while_the_next_one_runs(ctx):
    """BakeDriver waits per window (mmdgpu_frames_wait_downloads), not for the whole context: when the sink for window
    k-1 starts, window k has been issued and is normally still running / copying.  Results are checked against the
    oracle; the overlap through the non-blocking mmdgpu_frames_downloads_done of the OTHER frames object."""
    from simple_mmd_renderer_b200 import shard
    cfg, model, motion = synth_case("C1")
    m = Model(ctx, model)
    a = Motion(m, motion)
    window = 128
    drv = shard.BakeDriver(m, a, window)
    orc = _oracle(model, motion)
    seen, next_in_flight, order = [], [], []

    def sink(first, n, pos, nrm):
        k = len(seen)
        # window k+1 was issued before this call (except after the last one)
        order.append(drv.issued)
        other = drv.frames[(k + 1) & 1]
        next_in_flight.append(not other.downloads_done())
        ref = orc.run_frame(first + n - 1)
        assert_bitwise(pos[n - 1], ref["pos"], f"window at {first}: last frame positions")
        assert_bitwise(nrm[n - 1], ref["nrm"], f"window at {first}: last frame normals")
        seen.append((first, n))
    drv.run(10, 10 + 5 * window + 17, sink)
    assert seen == [(10 + i * window, window) for i in range(5)] + [(10 + 5 * window, 17)]
    assert order[:-1] == [k + 2 for k in range(len(order) - 1)], "window k+1 must be issued before window k is delivered"
    # 128 frames x 50 k vertices: 150 MB of copies per window (~3 ms) against microseconds of host work before the check
    assert any(next_in_flight[:-1]), "every next window had already finished when its predecessor was delivered: no overlap"
    drv.close()


def test_bake_driver_overlaps_a_slow_sink_with_the_device(ctx):
    """Wall-clock form of the same property: a sink that takes as long as a window takes on the device must cost
    about max(sink, device) per window, not their sum."""
    from simple_mmd_renderer_b200 import shard
    cfg, model, motion = synth_case("C1")
    m = Model(ctx, model)
    a = Motion(m, motion)
    drv = shard.BakeDriver(m, a, 128)
    n_win = 8

    def run(delay):
        t0 = time.perf_counter()
        drv.run(0, n_win * 128, lambda *_: time.sleep(delay))
        ctx.synchronize()
        return time.perf_counter() - t0
    run(0.0)
    base = min(run(0.0) for _ in range(3))          # device-bound: ~ n_win x (evaluate + copy)
    per = base / n_win
    slow = min(run(per) for _ in range(3))
    assert slow < 1.6 * base, f"sink and device ran one after the other: {slow:.4f} s against {base:.4f} s without a sink delay"
    drv.close()


@pytest.mark.parametrize("n_slots", [1, 45, 300])
def test_nested_ik_solves(ctx, n_slots):
    """synth.make_ik_nested(): an IK bone as the target of another solve (its solve runs after every CCD step of the
    outer one), an IK bone as a link (its solve runs when the outer solve first re-evaluates its links) and three levels
    deep — libmmd's recursion through UpdateBoneTransform (poser_impl.inl:203-206, :303).  Fused and step-wise, small
    and large batches (the large one would take the chain-local-image kernel, which nested models must not)."""
    _, model, motion = synth_case("ik_nested")
    orc = _oracle(model, motion)
    m = Model(ctx, model)
    a = Motion(m, motion)
    frames = [(7 * k) % 45 for k in range(n_slots)]
    fr = Frames(m, 1, n_slots)
    fr.update(a, frames)
    for k in sorted(set([0, n_slots // 2, n_slots - 1] + list(range(min(n_slots, 45))))):
        ref = orc.run_frame(frames[k])
        assert_bitwise(fr.bone_local_matrices(k), ref["local"], f"slot {k} frame {frames[k]} local matrices")
        assert_bitwise(fr.bone_matrices(k), ref["skin"], f"slot {k} frame {frames[k]} skinning matrices")
        assert_bitwise(fr.download(k, capi.STREAM_POSITION), ref["pos"], f"slot {k} frame {frames[k]} positions")
        assert_bitwise(fr.download(k, capi.STREAM_NORMAL), ref["nrm"], f"slot {k} frame {frames[k]} normals")
    # libmmd's call sequence, one frame at a time
    one = Frames(m, 1, 1)
    for f in (3, 22, 40):
        one.reset_posing(); one.seek_frame(a, [f]); one.pre_physics_posing(); one.post_physics_posing(); one.deform()
        ref = orc.run_frame(f)
        assert_bitwise(one.bone_matrices(0), ref["skin"], f"step-wise frame {f} skinning matrices")
        assert_bitwise(one.download(0, capi.STREAM_POSITION), ref["pos"], f"step-wise frame {f} positions")


def test_reset_and_seek_is_reset_then_seek(ctx):
    """mmdgpu_reset_and_seek_frame / _time = reset_posing + seek in one sampling launch: untracked bones / morphs get
    identity / zero, tracked ones their key frames - also after manual poses and a different clip had been set."""
    cfg, model, motion = synth_case("tiny_full")
    other = synth.make_motion(cfg, model, instance=3)
    m = Model(ctx, model)
    a, b = Motion(m, motion), Motion(m, other)
    two, one = Frames(m, 1, 3), Frames(m, 1, 3)
    frames = [4, 55, 89]
    q = np.asarray([0.5, 0.5, 0.5, 0.5], np.float32)
    for fr in (two, one):
        fr.reset_posing(); fr.seek_frame(b, [9, 9, 9])
        for s in range(3):
            fr.set_bone_pose(s, 1, [1, 2, 3], q)
            fr.set_morph_pose(s, 0, 0.7)
    two.reset_posing(); two.seek_frame(a, frames)
    one.reset_and_seek_frame(a, frames)
    for fr in (two, one):
        fr.pre_physics_posing(); fr.post_physics_posing(); fr.deform()
    for k in range(3):
        assert_bitwise(one.bone_poses(k), two.bone_poses(k), f"slot {k} poses")
        assert_bitwise(one.morph_rates(k), two.morph_rates(k), f"slot {k} rates")
        assert_bitwise(one.download(k, capi.STREAM_POSITION), two.download(k, capi.STREAM_POSITION), f"slot {k} positions")
    times = [0.25, 1.0123, 2.9]
    two.reset_posing(); two.seek_time(a, times)
    one.reset_and_seek_time(a, times)
    for k in range(3):
        assert_bitwise(one.bone_poses(k), two.bone_poses(k), f"slot {k} poses (time)")
    # single-slot objects pass the frame id / time as a kernel argument: same results as the array path
    s1 = Frames(m, 1, 1)
    orc = _oracle(model, motion)
    for f in (0, 17, 90):
        s1.update(a, [f])
        assert_bitwise(s1.download(0, capi.STREAM_POSITION), orc.run_frame(f)["pos"], f"by-value frame {f}")
    s1.reset_and_seek_time(a, [1.5]); s1.pre_physics_posing(); s1.post_physics_posing(); s1.deform()
    assert_bitwise(s1.download(0, capi.STREAM_POSITION), orc.run_time(1.5)["pos"], "by-value time 1.5 s")


def test_cxx_shim_replays_recorded_calls_exactly(ctx, tmp_path):
    """include/mmdgpu.hpp records ResetPosing / SeekFrame / PrePhysicsPosing and fuses main.cpp's sequence; every other
    continuation must behave as if each call had been issued when it was made (tests/cxx/shim_sequences.cc)."""
    import subprocess
    import pmxio
    from conftest import ROOT
    from simple_mmd_renderer_b200 import lib
    cfg, model, motion = synth_case("tiny_full")          # has post-physics bones, IK, bone / group morphs
    (tmp_path / "m.pmx").write_bytes(pmxio.write_pmx(model, version=2.1))      # QDEF-tagged vertices need a 2.1 container
    (tmp_path / "m.vmd").write_bytes(pmxio.write_vmd(motion))
    exe = tmp_path / "shim_sequences"
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", f"-I{ROOT}/include", f"{ROOT}/tests/cxx/shim_sequences.cc",
                        lib.SO_PATH, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(tmp_path / "m.pmx"), str(tmp_path / "m.vmd")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("same") == 12 and "DIFFERENT" not in r.stdout, r.stdout


@pytest.mark.parametrize("name", ["tiny_full", "ik_zoo", "C2"])
def test_pose_frame_is_the_four_libmmd_calls(ctx, name):
    """mmdgpu_pose_frame / _time = ResetPosing; SeekFrame / SeekTime; PrePhysicsPosing; PostPhysicsPosing in one call (one
    sampling launch, one hierarchy pass): same poses, rates, matrices and vertices as the four separate calls, and as
    libmmd."""
    _, model, motion = synth_case(name)
    m = Model(ctx, model)
    a = Motion(m, motion)
    n = 5
    frames = [0, 3, 17, 29, 40]
    four, one = Frames(m, 1, n), Frames(m, 1, n)
    for fr in (four, one):                                  # leftovers a reset must wipe
        fr.update(a, [9] * n)
        fr.set_morph_pose(2, 0, 0.5) if m.n_morphs else None
        fr.set_bone_pose(1, 1, [1, 2, 3], [0.5, 0.5, 0.5, 0.5])
    four.reset_posing(); four.seek_frame(a, frames); four.pre_physics_posing(); four.post_physics_posing(); four.deform()
    one.pose_frame(a, frames); one.deform()
    orc = _oracle(model, motion)
    for k, f in enumerate(frames):
        ref = orc.run_frame(f)
        assert_bitwise(one.bone_poses(k), four.bone_poses(k), f"{name} slot {k} poses")
        assert_bitwise(one.bone_matrices(k), four.bone_matrices(k), f"{name} slot {k} skinning matrices")
        assert_bitwise(one.bone_matrices(k), ref["skin"], f"{name} frame {f} skinning matrices vs oracle")
        assert_bitwise(one.download(k, capi.STREAM_POSITION), ref["pos"], f"{name} frame {f} positions vs oracle")
    times = [0.0, 0.31, 0.777, 1.0, 1.29]
    four.reset_posing(); four.seek_time(a, times); four.pre_physics_posing(); four.post_physics_posing(); four.deform()
    one.pose_time(a, times); one.deform()
    for k, t in enumerate(times):
        assert_bitwise(one.download(k, capi.STREAM_POSITION), orc.run_time(t)["pos"], f"{name} time {t} positions vs oracle")
        assert_bitwise(one.bone_matrices(k), four.bone_matrices(k), f"{name} time {t} skinning matrices")
    # one-slot object: frame id / time by kernel argument
    s1 = Frames(m, 1, 1)
    s1.pose_frame(a, [17]); s1.deform()
    assert_bitwise(s1.download(0, capi.STREAM_POSITION), orc.run_frame(17)["pos"], f"{name} one-slot pose_frame")
    s1.pose_time(a, [0.777]); s1.deform()
    assert_bitwise(s1.download(0, capi.STREAM_POSITION), orc.run_time(0.777)["pos"], f"{name} one-slot pose_time")
