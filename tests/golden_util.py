"""Helpers shared by the golden-fixture tests (CPU oracle and GPU)."""
import hashlib
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


def check_against_golden(g, f: int, got: dict, what: str):
    """`got` has pos / nrm / skin (+ optionally local / poses / rates): full-array SHA-256 plus the stored samples."""
    stride = int(g["stride"])
    np.testing.assert_array_equal(got["skin"].view(np.uint32), g[f"f{f}_skin"].view(np.uint32), err_msg=f"{what} skin")
    np.testing.assert_array_equal(got["pos"][::stride].view(np.uint32), g[f"f{f}_pos_s"].view(np.uint32),
                                  err_msg=f"{what} position sample")
    np.testing.assert_array_equal(got["nrm"][::stride].view(np.uint32), g[f"f{f}_nrm_s"].view(np.uint32),
                                  err_msg=f"{what} normal sample")
    for k in ("pos", "nrm", "skin", "local", "poses", "rates"):
        if k in got:
            assert sha(got[k]) == str(g[f"f{f}_{k}_sha"]), f"{what}: SHA-256 of {k} differs from libmmd's"
