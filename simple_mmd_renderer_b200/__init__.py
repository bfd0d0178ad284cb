"""mmdgpu — the libmmd per-frame deformation path of simple_mmd_renderer on B200.

The product is `libmmdgpu.so` (hand-written CUDA for sm_100a behind the C-ABI of include/mmdgpu.h, sources in
`csrc/`).  This package is the thin Python side used by the tests and bench.py:

  lib      ctypes loader of the shared library (raises if it is missing: there is no CPU fallback)
  capi     ctypes mirror of the descriptor structs and enums of mmdgpu.h
  poser    Context / Model / Motion / Frames and the libmmd-named Poser / MotionPlayer
  shard    multi-GPU decomposition (instances of a crowd, frame ranges of a bake) and the windowed gather
  synth    deterministic synthetic PMX / VMD-shaped inputs of the BASELINE configs
"""
__all__ = ["lib", "capi", "poser", "shard", "synth"]
