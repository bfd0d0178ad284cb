"""TEST INFRASTRUCTURE — CPU checkers for the deformation path.  Not product code.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.  The
product (simple_mmd_renderer_b200/) never does, and fails loudly when its CUDA library is missing.

Two checkers, same flat descriptors (include/mmdgpu.h) in, same arrays out:

* ``Reference`` — libmmd itself, compiled by oracle/Makefile from the headers under
  /root/reference into oracle/_ref/libmmd_ref.so (git-ignored; travels to the GPU box as a built file).
* ``Restatement`` — oracle/mmd_oracle.c, a plain-C restatement of the algorithm,
  built into oracle/libmmd_oracle.so.  It is pinned bit-for-bit against ``Reference`` by
  tests/test_oracle_pin.py and against the committed fixtures under tests/golden/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from simple_mmd_renderer_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libmmd_ref.so")
REF_FAST_SO = os.path.join(HERE, "_ref", "libmmd_ref_fast.so")
PORT_SO = os.path.join(HERE, "libmmd_oracle.so")


def build(verbose: bool = False) -> None:
    """Compile both checkers (the reference one only where /root/reference is mounted)."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _fp(a, dtype=np.float32):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _Session:
    """Common driver over the identical C interfaces of the two checkers."""

    prefix = ""
    so_path = ""

    def __init__(self, model: dict, motion: dict | None):
        if not os.path.exists(self.so_path):
            raise FileNotFoundError(f"{self.so_path} not built (run `make -C oracle`)")
        self.lib = C.CDLL(self.so_path)
        p = self.prefix
        self._create = getattr(self.lib, p + "create")
        self._create.restype = C.c_void_p
        self._create.argtypes = [C.c_void_p, C.c_void_p]
        self._destroy = getattr(self.lib, p + "destroy")
        self._destroy.argtypes = [C.c_void_p]
        self._run = getattr(self.lib, p + "run_frame")
        self._run.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 6
        self._manual = getattr(self.lib, p + "run_manual")
        self._manual.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self._run_time = getattr(self.lib, p + "run_time")
        self._run_time.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 5
        self._override = getattr(self.lib, p + "run_frame_override")
        self._override.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 7
        self._skin = getattr(self.lib, p + "get_skinning")
        self._skin.argtypes = [C.c_void_p] * 4
        self._ik = getattr(self.lib, p + "get_ik_class")
        self._ik.restype = C.c_uint32
        self._ik.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        self._repack = getattr(self.lib, p + "repack_sokol32")
        self._repack.argtypes = [C.c_void_p, C.c_void_p]
        self._time = getattr(self.lib, p + "time_frames")
        self._time.restype = C.c_double
        self._time.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        self.model = model
        self.nv, self.nb, self.nm = int(model["n_vertices"]), int(model["n_bones"]), int(model["n_morphs"])
        md, self._keep_m = capi.model_desc(model)
        if motion is not None:
            ad, self._keep_a = capi.anim_desc(motion)
            self.h = self._create(C.byref(md), C.byref(ad))
        else:
            self.h = self._create(C.byref(md), None)
        if not self.h:
            raise RuntimeError("oracle session creation failed")

    def close(self):
        if getattr(self, "h", None):
            self._destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_frame(self, frame: int, want=("pos", "nrm", "skin", "local", "poses", "rates")) -> dict:
        out = {}
        shapes = dict(pos=(self.nv, 3), nrm=(self.nv, 3), skin=(self.nb, 16), local=(self.nb, 16),
                      poses=(self.nb, 7), rates=(self.nm,))
        ptrs = []
        for k in ("pos", "nrm", "skin", "local", "poses", "rates"):
            if k in want:
                out[k] = np.zeros(shapes[k], np.float32)
                ptrs.append(_fp(out[k]))
            else:
                ptrs.append(None)
        self._run(self.h, int(frame), *ptrs)
        return out

    def run_time(self, seconds: float) -> dict:
        """ResetPosing; MotionPlayer::SeekTime(seconds); Pre; Post; Deform."""
        out = dict(pos=np.zeros((self.nv, 3), np.float32), nrm=np.zeros((self.nv, 3), np.float32),
                   skin=np.zeros((self.nb, 16), np.float32), poses=np.zeros((self.nb, 7), np.float32),
                   rates=np.zeros((self.nm,), np.float32))
        self._run_time(self.h, float(seconds), _fp(out["pos"]), _fp(out["nrm"]), _fp(out["skin"]), _fp(out["poses"]),
                       _fp(out["rates"]))
        return out

    def run_frame_override(self, frame: int, bones, skin16, local16=None) -> dict:
        """ResetPosing; SeekFrame; Pre; overwrite the listed bones' skinning (and local) matrices as a host physics
        reactor would (mmd-bullet_impl.inl:34-56); Post; Deform."""
        bones = np.ascontiguousarray(bones, np.int32)
        skin16 = np.ascontiguousarray(skin16, np.float32).reshape(bones.size, 16)
        loc = None if local16 is None else np.ascontiguousarray(local16, np.float32).reshape(bones.size, 16)
        out = dict(pos=np.zeros((self.nv, 3), np.float32), nrm=np.zeros((self.nv, 3), np.float32),
                   skin=np.zeros((self.nb, 16), np.float32), local=np.zeros((self.nb, 16), np.float32))
        self._override(self.h, int(frame), bones.size, _fp(bones), _fp(skin16), _fp(loc), _fp(out["pos"]), _fp(out["nrm"]),
                       _fp(out["skin"]), _fp(out["local"]))
        return out

    def run_manual(self, bones, poses7, morphs, weights) -> dict:
        bones = np.ascontiguousarray(bones, np.int32)
        poses7 = np.ascontiguousarray(poses7, np.float32)
        morphs = np.ascontiguousarray(morphs, np.int32)
        weights = np.ascontiguousarray(weights, np.float32)
        out = dict(pos=np.zeros((self.nv, 3), np.float32), nrm=np.zeros((self.nv, 3), np.float32),
                   skin=np.zeros((self.nb, 16), np.float32))
        self._manual(self.h, bones.size, _fp(bones), _fp(poses7), morphs.size, _fp(morphs), _fp(weights),
                     _fp(out["pos"]), _fp(out["nrm"]), _fp(out["skin"]))
        return out

    def skinning(self):
        t = np.zeros(self.nv, np.uint8)
        ids = np.zeros((self.nv, 4), np.int32)
        w = np.zeros((self.nv, 4), np.float32)
        self._skin(self.h, _fp(t), _fp(ids), _fp(w))
        return t, ids, w

    def ik_class(self):
        n = int(self.model["n_ik_links"])
        fix = np.zeros(max(n, 1), np.uint8)
        order = np.zeros(max(n, 1), np.uint8)
        k = self._ik(self.h, _fp(fix), _fp(order), n)
        return fix[:k], order[:k]

    def repack_sokol32(self):
        out = np.zeros((self.nv, 8), np.float32)
        self._repack(self.h, _fp(out))
        return out

    def time_frames(self, frames, n_threads: int = 1):
        """Wall seconds for the main.cpp:1788-1821 loop over `frames` on n_threads host threads."""
        frames = np.ascontiguousarray(frames, np.uint32)
        ck = C.c_double(0)
        sec = self._time(self.h, _fp(frames), frames.size, int(n_threads), C.byref(ck))
        return float(sec), float(ck.value)


class Reference(_Session):
    """libmmd itself (oracle/_ref/libmmd_ref.so)."""
    prefix = "ref_"
    so_path = REF_SO


class ReferenceFast(_Session):
    """libmmd at -O3 -march=x86-64-v3 with contraction on: timing only (bench.py), never a checker."""
    prefix = "ref_"
    so_path = REF_FAST_SO


class Restatement(_Session):
    """Plain-C restatement (oracle/libmmd_oracle.so)."""
    prefix = "port_"
    so_path = PORT_SO


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def have_reference_fast() -> bool:
    """Built, and this host has the AVX2 + FMA the fast build assumes."""
    if not os.path.exists(REF_FAST_SO):
        return False
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " avx2" in flags and " fma" in flags and " bmi2" in flags


def have_restatement() -> bool:
    return os.path.exists(PORT_SO)
