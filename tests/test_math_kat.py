"""Function-level known-answer tests: libmmd's math functions - Bezier lookup, NLerp, SLerp, the
three Euler conversions both ways, AxisToQuaternion, quaternion product / rotation matrix / inverse, the 4 x 4 product,
Normalize - one function at a time, bit for bit.  The answers in tests/golden/math_kat.npz were computed by libmmd itself
(tests/golden/make_math_kat.py); the CPU test holds the C restatement to them (and libmmd again where its harness is
built), the GPU test holds the device functions of csrc/mmd_math.cuh to them through the mmdgpu_test_math export."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
import oracle

KIN = [5, 9, 9, 5, 4, 4, 8, 4, 4, 32, 3]
KOUT = [1, 4, 4, 3, 4, 4, 4, 9, 4, 16, 3]
NAMES = ["bezier", "nlerp", "slerp", "quat_to_euler", "euler_to_quat", "axis_to_quat", "quat_mul", "quat_to_rows",
         "quat_inverse", "mat_mul", "vec_normalize"]


def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "math_kat.npz"))


def same_bits(got, want, what):
    """Bit-exact, except that any NaN equals any NaN (x86 and the GPU produce different default NaN payloads)."""
    g, w = got.view(np.uint32), want.view(np.uint32)
    bad = (g != w) & ~(np.isnan(got) & np.isnan(want))
    if bad.any():
        i = tuple(np.argwhere(bad)[0])
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.size} floats differ; first at {i}: {got[i]!r} vs {want[i]!r}")


def _run(lib, fn, op, x):
    out = np.zeros((x.shape[0], KOUT[op]), np.float32)
    f = getattr(lib, fn)
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
    assert f(op, x.ctypes.data_as(C.c_void_p), x.shape[0], out.ctypes.data_as(C.c_void_p)) == 0
    return out


def _affine12(m):
    return np.ascontiguousarray(m.reshape(-1, 4, 4)[:, :, :3])


@pytest.mark.parametrize("op", range(11), ids=NAMES)
def test_restatement_reproduces_libmmd_known_answers(op):
    g = golden()
    x, want = np.ascontiguousarray(g[f"in_{op}"]), g[f"out_{op}"]
    assert x.shape[1] == KIN[op] and want.shape[1] == KOUT[op] and x.shape[0] >= 400
    got = _run(C.CDLL(oracle.PORT_SO), "port_math_kat", op, x)
    same_bits(got, want, f"restatement {NAMES[op]}")
    if oracle.have_reference():          # the fixture is what libmmd computes today
        same_bits(_run(C.CDLL(oracle.REF_SO), "ref_math_kat", op, x), want, f"libmmd {NAMES[op]}")


def test_known_answers_cover_the_branches():
    """The fixture must actually reach the cut-offs it claims to: NLerp / SLerp shortcuts, NaN from unclamped asin / acos,
    the identity shortcut of AxisToQuaternion, linear Bezier curves and the last table node."""
    g = golden()
    nl_in, nl_out = g["in_1"], g["out_1"]
    assert (nl_out == nl_in[:, :4]).all(1).any() and (nl_out == nl_in[:, 4:8]).all(1).any()
    assert np.isnan(g["out_3"]).any() or np.isnan(g["out_2"]).any()
    ax = g["out_5"]
    assert ((ax == np.asarray([0, 0, 0, 1], np.float32)).all(1)).sum() >= 2
    bz_in, bz_out = g["in_0"], g["out_0"]
    lin = (bz_in[:, 0] == bz_in[:, 1]) & (bz_in[:, 2] == bz_in[:, 3])
    assert lin.any() and (bz_out[lin, 0] == bz_in[lin, 4]).all()
    assert (bz_in[:, 4] == 1.0).any()


@pytest.mark.gpu
@pytest.mark.parametrize("op", range(11), ids=NAMES)
def test_device_functions_reproduce_libmmd_known_answers(ctx, op):
    from simple_mmd_renderer_b200.lib import check
    g = golden()
    x, want = np.ascontiguousarray(g[f"in_{op}"]), g[f"out_{op}"]
    got = np.zeros_like(want)
    check(ctx.lib.mmdgpu_test_math(ctx.h, op, x.ctypes.data_as(C.c_void_p), x.shape[0], got.ctypes.data_as(C.c_void_p)), ctx.h)
    if op == 9:     # the device carries the 12 affine elements of a matrix; the fourth column is implied
        got, want = _affine12(got), _affine12(want)
    same_bits(got, want, f"device {NAMES[op]}")
