"""Pins the CPU restatement (oracle/mmd_oracle.c) against the reference: bit-for-bit against libmmd itself where
the reference harness could be built (this container), and against the committed libmmd-generated fixtures
(tests/golden/*.npz, made by tests/golden/make_golden.py) everywhere."""
import numpy as np
import pytest

import oracle
from conftest import assert_bitwise, synth_case
from golden_util import check_against_golden, input_digest, load_golden, sha

CASES = ["tiny", "tiny_full", "small", "ik_zoo", "ik_nested"]


@pytest.mark.parametrize("name", CASES)
def test_generator_is_deterministic(name):
    cfg, model, motion = synth_case(name)
    g = load_golden(name)
    assert input_digest(model, motion) == str(g["input_digest"]), \
        "synthetic generator drifted from the inputs the golden fixtures were made with"


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_golden(name):
    """ik_zoo covers what the random configs do not: FIX_Y / FIX_Z / FIX_ALL links, the YZX Euler order, 3- and
    4-link chains, swapped limits, the 256-iteration cap, an IK bone that sorts before its links."""
    cfg, model, motion = synth_case(name)
    g = load_golden(name)
    port = oracle.Restatement(model, motion)
    t, ids, w = port.skinning()
    np.testing.assert_array_equal(t, g["norm_type"])
    np.testing.assert_array_equal(ids, g["norm_ids"])
    np.testing.assert_array_equal(w.view(np.uint32), g["norm_w"].view(np.uint32))
    fix, order = port.ik_class()
    np.testing.assert_array_equal(fix, g["ik_fix"])
    np.testing.assert_array_equal(order, g["ik_order"])
    for f in g["frames"]:
        f = int(f)
        got = port.run_frame(f)
        check_against_golden(g, f, got, f"{name} frame {f}")
        assert sha(port.repack_sokol32()) == str(g[f"f{f}_sokol32_sha"])


@pytest.mark.skipif(not oracle.have_reference(), reason="libmmd reference harness not built (needs /root/reference)")
@pytest.mark.parametrize("name,frames", [("tiny", range(0, 70, 3)), ("tiny_full", range(0, 100)), ("small", [0, 7, 60, 119]),
                                         ("ik_zoo", range(0, 45)), ("ik_nested", range(0, 45))])
def test_restatement_matches_libmmd_bitwise(name, frames):
    cfg, model, motion = synth_case(name)
    ref = oracle.Reference(model, motion)
    port = oracle.Restatement(model, motion)
    for a, b in zip(ref.skinning(), port.skinning()):
        np.testing.assert_array_equal(a, b)
    for f in frames:
        r, p = ref.run_frame(f), port.run_frame(f)
        for k in ("poses", "rates", "local", "skin", "pos", "nrm"):
            assert_bitwise(p[k], r[k], f"{name} frame {f} {k}")
        assert_bitwise(port.repack_sokol32(), ref.repack_sokol32(), f"{name} frame {f} sokol32")


@pytest.mark.skipif(not oracle.have_reference(), reason="libmmd reference harness not built (needs /root/reference)")
@pytest.mark.parametrize("name,frames", [("C1", [0, 7, 150, 299]), ("C2", [0, 7, 33, 150, 299]), ("C3", [3, 144, 167, 357, 399])])
def test_restatement_matches_libmmd_on_baseline_configs(name, frames):
    """The BASELINE configs at full size, at every frame tests/test_gpu_parity.py evaluates on them: wherever a GPU test
    falls back to the restatement, it has been compared with libmmd on exactly that input."""
    cfg, model, motion = synth_case(name)
    ref = oracle.Reference(model, motion)
    port = oracle.Restatement(model, motion)
    for f in frames:
        r, p = ref.run_frame(f), port.run_frame(f)
        for k in ("poses", "rates", "local", "skin", "pos", "nrm"):
            assert_bitwise(p[k], r[k], f"{name} frame {f} {k}")


@pytest.mark.skipif(not oracle.have_reference(), reason="libmmd reference harness not built (needs /root/reference)")
def test_restatement_manual_posing_matches_libmmd():
    cfg, model, motion = synth_case("tiny_full")
    rng = np.random.default_rng(3)
    nb, nm = model["n_bones"], model["n_morphs"]
    bones = rng.choice(nb, 10, replace=False).astype(np.int32)
    q = rng.normal(size=(10, 4)).astype(np.float32)
    q /= np.sqrt((q * q).sum(1, keepdims=True)).astype(np.float32)
    pose7 = np.concatenate([rng.uniform(-1, 1, (10, 3)).astype(np.float32), q], 1)
    morphs = np.arange(nm, dtype=np.int32)
    w = rng.uniform(-0.3, 1.0, nm).astype(np.float32)
    r = oracle.Reference(model, None).run_manual(bones, pose7, morphs, w)
    p = oracle.Restatement(model, None).run_manual(bones, pose7, morphs, w)
    for k in ("skin", "pos", "nrm"):
        assert_bitwise(p[k], r[k], f"manual {k}")


@pytest.mark.skipif(not oracle.have_reference(), reason="libmmd reference harness not built (needs /root/reference)")
@pytest.mark.parametrize("with_local", [True, False])
def test_restatement_physics_hand_back_matches_libmmd(with_local):
    """Skinning / local matrices overwritten between Pre and Post, as a host physics reactor does
    (mmd-bullet_impl.inl:34-56): post-physics children and Deform must see them the same way in both checkers."""
    cfg, model, motion = synth_case("tiny_full")
    rng = np.random.default_rng(8)
    pre = np.flatnonzero((model["bone_flags"] & 0x1000) == 0)
    bones = rng.choice(pre, 8, replace=False).astype(np.int32)
    skin = rng.normal(size=(8, 16)).astype(np.float32)
    local = rng.normal(size=(8, 16)).astype(np.float32) if with_local else None
    for f in (0, 14, 41):
        r = oracle.Reference(model, motion).run_frame_override(f, bones, skin, local)
        p = oracle.Restatement(model, motion).run_frame_override(f, bones, skin, local)
        for k in r:
            assert_bitwise(p[k], r[k], f"hand-back frame {f} {k}")


@pytest.mark.parametrize("name", CASES)
def test_restatement_seek_time_matches_golden(name):
    """MotionPlayer::SeekTime(double): sub-frame sampling, barycentre in double, no key-frame snapping."""
    cfg, model, motion = synth_case(name)
    g = load_golden(name)
    port = oracle.Restatement(model, motion)
    for i, t in enumerate(g["times"]):
        got = port.run_time(float(t))
        np.testing.assert_array_equal(got["poses"].view(np.uint32), g[f"t{i}_poses"].view(np.uint32))
        for k in ("pos", "nrm", "skin", "poses", "rates"):
            assert sha(got[k]) == str(g[f"t{i}_{k}_sha"]), f"{name} t={t}: {k} differs from libmmd's"


@pytest.mark.skipif(not oracle.have_reference(), reason="libmmd reference harness not built (needs /root/reference)")
def test_restatement_seek_time_matches_libmmd_bitwise():
    cfg, model, motion = synth_case("tiny_full")
    ref = oracle.Reference(model, motion)
    port = oracle.Restatement(model, motion)
    rng = np.random.default_rng(9)
    for t in list(rng.uniform(0, 3.2, 40)) + [0.0, 1 / 30, 2 / 30, 1.0, 3.0, 100.0]:
        r, p = ref.run_time(float(t)), port.run_time(float(t))
        for k in ("poses", "rates", "skin", "pos", "nrm"):
            assert_bitwise(p[k], r[k], f"t={t} {k}")


def test_multithreaded_timing_entry_is_consistent():
    """port_time_frames (the CPU-baseline leg of bench.py): same checksum for 1 and 4 threads."""
    cfg, model, motion = synth_case("tiny")
    port = oracle.Restatement(model, motion)
    frames = np.arange(40, dtype=np.uint32)
    _, c1 = port.time_frames(frames, 1)
    _, c4 = port.time_frames(frames, 4)
    assert abs(c1 - c4) <= 1e-6 * max(1.0, abs(c1))
