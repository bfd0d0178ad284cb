"""Loader for libmmdgpu.so (the C-ABI of include/mmdgpu.h).

The library is built in-tree by `simple_mmd_renderer_b200/csrc/Makefile` (see `build_library`).  There is no
fallback of any kind: if the shared object is missing, `load()` raises; if no CUDA device is present, the
first call that needs one returns MMDGPU_ERR_CUDA and the wrappers raise `MmdGpuError`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libmmdgpu.so")
CSRC = os.path.join(HERE, "csrc")

_lib = None


class MmdGpuError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"mmdgpu status {status}: {message}")
        self.status = status
        self.message = message


def build_library(verbose: bool = False, jobs: int = 4) -> str:
    """Compile libmmdgpu.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, f"-j{jobs}", "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-6000:], r.stderr[-6000:])
    if r.returncode != 0:
        raise RuntimeError("libmmdgpu.so build failed")
    return SO_PATH


_u32, _i32, _sz, _vp, _f32 = C.c_uint32, C.c_int32, C.c_size_t, C.c_void_p, C.c_float
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors every MMDGPU_API declaration of include/mmdgpu.h
SIGNATURES = {
    "mmdgpu_version": (C.c_int, []),
    "mmdgpu_status_string": (C.c_char_p, [C.c_int]),
    "mmdgpu_context_create": (C.c_int, [C.c_int, _vp, _PP]),
    "mmdgpu_context_destroy": (None, [_vp]),
    "mmdgpu_last_error": (C.c_char_p, [_vp]),
    "mmdgpu_context_synchronize": (C.c_int, [_vp]),
    "mmdgpu_context_stream": (_vp, [_vp]),
    "mmdgpu_context_launch_count": (C.c_uint64, [_vp]),
    "mmdgpu_context_set_profiling": (C.c_int, [_vp, C.c_int]),
    "mmdgpu_context_profile_read": (C.c_int, [_vp, _vp, _vp]),
    "mmdgpu_context_join_downloads": (C.c_int, [_vp]),
    "mmdgpu_model_create_from_arrays": (C.c_int, [_vp, _vp, _vp, _PP]),
    "mmdgpu_model_create_from_pmx": (C.c_int, [_vp, _vp, _sz, _vp, _PP]),
    "mmdgpu_model_destroy": (None, [_vp]),
    "mmdgpu_model_vertex_count": (_u32, [_vp]),
    "mmdgpu_model_bone_count": (_u32, [_vp]),
    "mmdgpu_model_morph_count": (_u32, [_vp]),
    "mmdgpu_model_material_count": (_u32, [_vp]),
    "mmdgpu_model_plan": (_vp, [_vp]),
    "mmdgpu_model_find_bone": (_i32, [_vp, _vp, _sz]),
    "mmdgpu_model_find_morph": (_i32, [_vp, _vp, _sz]),
    "mmdgpu_animation_create_from_arrays": (C.c_int, [_vp, _vp, _vp, _PP]),
    "mmdgpu_animation_create_from_vmd": (C.c_int, [_vp, _vp, _vp, _sz, _PP]),
    "mmdgpu_animation_destroy": (None, [_vp]),
    "mmdgpu_animation_length": (_u32, [_vp]),
    "mmdgpu_frames_create": (C.c_int, [_vp, _vp, _u32, _u32, C.c_int, _PP]),
    "mmdgpu_frames_destroy": (None, [_vp]),
    "mmdgpu_frames_slot_count": (_u32, [_vp]),
    "mmdgpu_reset_posing": (C.c_int, [_vp]),
    "mmdgpu_seek_frame": (C.c_int, [_vp, _vp, _vp]),
    "mmdgpu_seek_frame_range": (C.c_int, [_vp, _vp, _vp, _u32]),
    "mmdgpu_seek_time": (C.c_int, [_vp, _vp, _vp]),
    "mmdgpu_pose_frame": (C.c_int, [_vp, _vp, _vp]),
    "mmdgpu_pose_time": (C.c_int, [_vp, _vp, _vp]),
    "mmdgpu_reset_and_seek_frame": (C.c_int, [_vp, _vp, _vp]),
    "mmdgpu_reset_and_seek_time": (C.c_int, [_vp, _vp, _vp]),
    "mmdgpu_set_bone_pose": (C.c_int, [_vp, _u32, _u32, _vp, _vp]),
    "mmdgpu_set_morph_pose": (C.c_int, [_vp, _u32, _u32, _f32]),
    "mmdgpu_pre_physics_posing": (C.c_int, [_vp]),
    "mmdgpu_post_physics_posing": (C.c_int, [_vp]),
    "mmdgpu_deform": (C.c_int, [_vp]),
    "mmdgpu_set_skinning_matrix_override": (C.c_int, [_vp, _u32, _u32, _vp, _vp]),
    "mmdgpu_update": (C.c_int, [_vp, _vp, _vp]),
    "mmdgpu_update_range": (C.c_int, [_vp, _vp, _vp, _u32]),
    "mmdgpu_frames_device_ptr": (C.c_int, [_vp, C.c_int, _PP, C.POINTER(_sz)]),
    "mmdgpu_frames_download": (C.c_int, [_vp, _u32, C.c_int, _vp, _sz]),
    "mmdgpu_frames_download_async": (C.c_int, [_vp, _u32, _u32, C.c_int, _vp, _sz]),
    "mmdgpu_frames_download_pair_async": (C.c_int, [_vp, _u32, _vp, _sz]),
    "mmdgpu_frames_wait_downloads": (C.c_int, [_vp]),
    "mmdgpu_frames_wait_skinning": (C.c_int, [_vp]),
    "mmdgpu_frames_downloads_done": (C.c_int, [_vp]),
    "mmdgpu_frames_bind_output": (C.c_int, [_vp, C.c_int, _vp, _sz]),
    "mmdgpu_bone_matrices_download": (C.c_int, [_vp, _u32, _vp]),
    "mmdgpu_bone_local_matrices_download": (C.c_int, [_vp, _u32, _vp]),
    "mmdgpu_bone_poses_download": (C.c_int, [_vp, _u32, _vp]),
    "mmdgpu_morph_rates_download": (C.c_int, [_vp, _u32, _vp]),
    "mmdgpu_material_images_download": (C.c_int, [_vp, _u32, _vp]),
    "mmdgpu_frames_slot_run": (_u32, [_vp]),
    "mmdgpu_peer_buffer_create": (C.c_int, [_vp, _sz, _PP, _vp]),
    "mmdgpu_peer_buffer_open": (C.c_int, [_vp, _vp, _PP]),
    "mmdgpu_peer_buffer_release": (C.c_int, [_vp, _vp, C.c_int]),
    "mmdgpu_test_math": (C.c_int, [_vp, C.c_int, _vp, _u32, _vp]),
    "mmdgpu_host_alloc": (C.c_int, [_sz, _PP]),
    "mmdgpu_host_free": (None, [_vp]),
    "mmdgpu_plan_create": (C.c_int, [_vp, _vp, _PP, C.c_char_p, _sz]),
    "mmdgpu_plan_create_from_pmx": (C.c_int, [_vp, _sz, _vp, _PP, C.c_char_p, _sz]),
    "mmdgpu_plan_destroy": (None, [_vp]),
    "mmdgpu_anim_plan_create": (C.c_int, [_vp, _u32, _u32, _PP, C.c_char_p, _sz]),
    "mmdgpu_anim_plan_create_from_vmd": (C.c_int, [_vp, _vp, _sz, _PP, C.c_char_p, _sz]),
    "mmdgpu_anim_plan_destroy": (None, [_vp]),
    "mmdgpu_anim_plan_get": (C.c_int, [_vp, C.c_int, _PP, C.POINTER(_sz)]),
    "mmdgpu_plan_get": (C.c_int, [_vp, C.c_int, _PP, C.POINTER(_sz)]),
    "mmdgpu_bezier_table": (C.c_int, [_vp, _vp]),
}


def load():
    """Return the ctypes handle of libmmdgpu.so; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise FileNotFoundError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {CSRC}`.  There is no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, ctx=None):
    if status != capi.OK:
        lib = load()
        msg = lib.mmdgpu_last_error(ctx)
        raise MmdGpuError(status, (msg or b"").decode("utf-8", "replace")
                          or lib.mmdgpu_status_string(status).decode())
