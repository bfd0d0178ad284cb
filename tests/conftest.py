"""Shared fixtures.  `-m "not gpu"` runs on a machine without a GPU; `-m gpu` are the parity tests proper."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product library and the CPU checkers exist (both are built files, not in git)."""
    from simple_mmd_renderer_b200 import lib
    import oracle
    if not os.path.exists(lib.SO_PATH):
        lib.build_library()
    if not oracle.have_restatement():
        oracle.build()
    yield


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    from simple_mmd_renderer_b200.poser import Context
    c = Context(0)
    yield c
    c.close()


_cache = {}


def synth_case(name: str):
    """(config, model arrays, motion arrays) for a named synthetic config, cached per session."""
    from simple_mmd_renderer_b200 import synth
    if name not in _cache:
        if name == "ik_zoo":            # hand-built rig that exercises every CCD IK branch
            model, motion = synth.make_ik_zoo()
            _cache[name] = (None, model, motion)
            return _cache[name]
        if name == "ik_nested":         # solves whose links / targets have IK themselves
            model, motion = synth.make_ik_nested()
            _cache[name] = (None, model, motion)
            return _cache[name]
        cfg = synth.CONFIGS[name]
        model = synth.make_model(cfg)
        motion = synth.make_motion(cfg, model)
        _cache[name] = (cfg, model, motion)
    return _cache[name]


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bitwise(a, b, what=""):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    neq = bits(a) != bits(b)
    if neq.any():
        idx = np.argwhere(neq)
        first = tuple(idx[0])
        raise AssertionError(f"{what}: {int(neq.sum())} of {neq.size} floats differ bitwise; first at {first}: "
                             f"{a[first]!r} vs {b[first]!r}; max abs diff {np.nanmax(np.abs(a - b))}")
