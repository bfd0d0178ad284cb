"""Generate tests/golden/math_kat.npz: known answers of libmmd's own math functions (L/util/math_impl.inl), computed by
libmmd itself through oracle/_ref/libmmd_ref.so (ref_math_kat).  Run where /root/reference is mounted:

    python tests/golden/make_math_kat.py

Inputs are seeded and stored next to the outputs, so the fixture is self-contained.  Row layouts: oracle/mmd_oracle.c,
port_math_kat.  Edge cases are appended by hand: NLerp / SLerp at and around the 1e-7 cut-offs, antipodal quaternions,
identical quaternions (omega = 0), Euler conversions at gimbal lock, asin arguments just beyond 1 (libmmd does not clamp:
NaN), zero and tiny rotation axes, Bezier x = 0 / 1 / table nodes, linear and extreme control points."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

KIN = [5, 9, 9, 5, 4, 4, 8, 4, 4, 32, 3]
KOUT = [1, 4, 4, 3, 4, 4, 4, 9, 4, 16, 3]
NAMES = ["bezier", "nlerp", "slerp", "quat_to_euler", "euler_to_quat", "axis_to_quat", "quat_mul", "quat_to_rows",
         "quat_inverse", "mat_mul", "vec_normalize"]


def unit_quats(rng, n):
    q = rng.normal(size=(n, 4)).astype(np.float32)
    return (q / np.sqrt((q * q).sum(1, keepdims=True)).astype(np.float32)).astype(np.float32)


def affine(rng, n):
    m = np.zeros((n, 4, 4), np.float32)
    m[:, :, :3] = rng.uniform(-3, 3, (n, 4, 3)).astype(np.float32)
    m[:, 3, 3] = 1.0
    return m.reshape(n, 16)


def inputs(rng):
    n = 400
    f32 = np.float32
    out = {}
    # 0 Bezier
    ctrl = rng.integers(0, 128, (n, 4)).astype(f32)
    x = rng.random(n).astype(f32)
    rows = np.concatenate([ctrl, x[:, None]], 1)
    special = [[20, 20, 107, 107, 0.3], [0, 0, 127, 127, 0.5], [127, 0, 0, 127, 0.0], [127, 0, 0, 127, 1.0], [0, 127, 127, 0, 0.999999],
               [64, 10, 64, 120, 1.0 / 31.0], [64, 10, 64, 120, 30.0 / 31.0], [1, 126, 126, 1, 0.5], [30, 30, 100, 101, 0.25]]
    out[0] = np.concatenate([rows, np.asarray(special, f32)]).astype(f32)
    # 1 NLerp / 2 SLerp
    a, b = unit_quats(rng, n), unit_quats(rng, n)
    l = rng.random(n).astype(f32)
    rows = np.concatenate([a, b, l[:, None]], 1)
    q0 = unit_quats(rng, 8)
    ls = np.asarray([0.0, 5e-8, 1e-7, 1.5e-7, 1 - 1.5e-7, 1 - 5e-8, 1.0, 0.5], f32)
    sp = [np.concatenate([q0, -q0, ls[:, None]], 1), np.concatenate([q0, q0, ls[:, None]], 1),
          np.concatenate([q0, np.roll(q0, 1, 0), ls[:, None]], 1),
          np.concatenate([q0, q0 + f32(1e-4), ls[:, None]], 1)]     # nearly identical, not normalised: comega > 1 -> acos NaN
    both = np.concatenate([rows] + sp).astype(f32)
    out[1] = both
    out[2] = both.copy()
    # 3 quaternion -> Euler, all three orders; gimbal lock and |asin argument| slightly above 1
    q = unit_quats(rng, n)
    order = rng.integers(0, 3, n).astype(f32)
    rows = np.concatenate([q, order[:, None]], 1)
    s = f32(np.sqrt(0.5))
    lock = []
    for o in range(3):
        for qq in ([s, 0, 0, s], [0, s, 0, s], [0, 0, s, s], [s, 0, 0, -s], [0.5, 0.5, 0.5, 0.5], [0.70710684, 0, 0, 0.70710684], [0, 0, 0, 1]):
            lock.append(list(qq) + [o])
    out[3] = np.concatenate([rows, np.asarray(lock, f32)]).astype(f32)
    # 4 Euler -> quaternion
    e = rng.uniform(-np.pi, np.pi, (n, 3)).astype(f32)
    rows = np.concatenate([e, rng.integers(0, 3, n).astype(f32)[:, None]], 1)
    sp = [[0, 0, 0, o] for o in range(3)] + [[np.pi, -np.pi, np.pi / 2, o] for o in range(3)] + [[1e-8, -1e-8, 40.0, o] for o in range(3)]
    out[4] = np.concatenate([rows, np.asarray(sp, f32)]).astype(f32)
    # 5 AxisToQuaternion (axis need not be unit; below 1e-7 -> identity)
    ax = rng.normal(size=(n, 3)).astype(f32)
    ang = rng.uniform(-7, 7, n).astype(f32)
    rows = np.concatenate([ax, ang[:, None]], 1)
    sp = [[0, 0, 0, 1.0], [1e-8, 0, 0, 1.0], [1e-7, 1e-7, 1e-7, 2.0], [5e-8, 5e-8, 5e-8, 2.0], [1, 0, 0, 0.0], [0, 0, 2, np.pi], [0, -3, 0, 2 * np.pi]]
    out[5] = np.concatenate([rows, np.asarray(sp, f32)]).astype(f32)
    # 6 product, 7 rotation rows, 8 inverse
    out[6] = np.concatenate([unit_quats(rng, n), rng.normal(size=(n, 4)).astype(f32)], 1)
    out[7] = np.concatenate([unit_quats(rng, n), rng.normal(size=(50, 4)).astype(f32), np.asarray([[0, 0, 0, 1], [1, 0, 0, 0], [0, 0, 0, -1]], f32)])
    out[8] = np.concatenate([unit_quats(rng, n), rng.uniform(-2, 2, (50, 4)).astype(f32)])
    # 9 affine 4 x 4 product (fourth column 0 0 0 1, as every matrix on the path)
    out[9] = np.concatenate([affine(rng, n), affine(rng, n)], 1)
    # 10 Normalize
    out[10] = np.concatenate([rng.normal(size=(n, 3)).astype(f32), np.asarray([[3, 0, 0], [0, -1e-20, 0], [1e19, 1e19, 0], [1e-7, 1e-7, 1e-7]], f32)])
    return {k: np.ascontiguousarray(v, f32) for k, v in out.items()}


def run(lib, fn, op, x):
    n = x.shape[0]
    assert x.shape[1] == KIN[op]
    out = np.zeros((n, KOUT[op]), np.float32)
    f = getattr(lib, fn)
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
    assert f(op, x.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)) == 0
    return out


def main():
    oracle.build()
    assert oracle.have_reference(), "needs /root/reference"
    ref = C.CDLL(oracle.REF_SO)
    rng = np.random.Generator(np.random.PCG64(20261018))
    ins = inputs(rng)
    out = {}
    for op, x in ins.items():
        out[f"in_{op}"] = x
        out[f"out_{op}"] = run(ref, "ref_math_kat", op, x)
    path = os.path.join(HERE, "math_kat.npz")
    np.savez_compressed(path, **out)
    print("math_kat", os.path.getsize(path), "bytes,", sum(v.shape[0] for v in ins.values()), "cases")


if __name__ == "__main__":
    main()
