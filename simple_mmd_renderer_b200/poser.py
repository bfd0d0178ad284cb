"""Host-side mirror of the libmmd interface that simple_mmd_renderer drives (main.cpp:1786-1825), on top of
the `mmdgpu_*` C-ABI.

libmmd name                                   here
--------------------------------------------  ----------------------------------------------
mmd::Model (L/model/model.inl)                Model        (flat arrays or PMX bytes)
mmd::Motion (L/motion/motion.inl)             Motion       (flat arrays or VMD bytes, bound to a Model)
mmd::Poser (L/motion/poser.inl:15-45)         Poser        ResetPosing / SetBonePose / SetMorphPose /
                                                           PrePhysicsPosing / PostPhysicsPosing / Deform /
                                                           pose_image
mmd::MotionPlayer (L/motion/poser.inl:184-198) MotionPlayer SeekFrame

`Frames` is the batched form the GPU wants (instances x frames slots per launch) used by the crowd and bake
drivers; `Poser` is a one-slot `Frames`.  All computation happens in libmmdgpu.so on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .lib import MmdGpuError, check, load

__all__ = ["Context", "Model", "Motion", "Frames", "Poser", "MotionPlayer", "PoseImage", "plan_arrays", "HostPlan",
           "bezier_table", "MmdGpuError"]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One GPU + stream (mmdgpu_context_t).  `stream` may be a raw cudaStream_t (e.g. torch's current stream)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.mmdgpu_context_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = int(device)

    def synchronize(self):
        check(self.lib.mmdgpu_context_synchronize(self.h), self.h)

    def set_profiling(self, enabled: bool):
        check(self.lib.mmdgpu_context_set_profiling(self.h, 1 if enabled else 0), self.h)

    def profile_read(self):
        """(ms_total[3], launches[3]) per kernel (pose_sample, hierarchy, skin) since the last read."""
        ms = (C.c_double * 3)()
        n = (C.c_uint64 * 3)()
        check(self.lib.mmdgpu_context_profile_read(self.h, ms, n), self.h)
        return [float(x) for x in ms], [int(x) for x in n]

    def join_downloads(self):
        check(self.lib.mmdgpu_context_join_downloads(self.h), self.h)

    @property
    def stream(self) -> int:
        return int(self.lib.mmdgpu_context_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.mmdgpu_context_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mmdgpu_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """Immutable device image of a PMX model (mmd::Model + Model::Normalize + Poser::Poser precomputation)."""

    def __init__(self, ctx: Context, arrays: dict | None = None, pmx_bytes: bytes | None = None,
                 extensions: bool = False):
        self.ctx = ctx
        lib = ctx.lib
        opt = capi.Options()
        opt.extensions = 1 if extensions else 0
        h = C.c_void_p()
        if arrays is not None:
            desc, keep = capi.model_desc(arrays)
            check(lib.mmdgpu_model_create_from_arrays(ctx.h, C.byref(desc), C.byref(opt), C.byref(h)), ctx.h)
            del keep
        elif pmx_bytes is not None:
            buf = (C.c_char * len(pmx_bytes)).from_buffer_copy(pmx_bytes)
            check(lib.mmdgpu_model_create_from_pmx(ctx.h, buf, len(pmx_bytes), C.byref(opt), C.byref(h)), ctx.h)
        else:
            raise ValueError("arrays or pmx_bytes required")
        self.h = h
        self.n_vertices = int(lib.mmdgpu_model_vertex_count(h))
        self.n_bones = int(lib.mmdgpu_model_bone_count(h))
        self.n_morphs = int(lib.mmdgpu_model_morph_count(h))
        self.n_materials = int(lib.mmdgpu_model_material_count(h))

    def plan(self) -> dict:
        return _plan_to_dict(self.ctx.lib, self.ctx.lib.mmdgpu_model_plan(self.h))

    def find_bone(self, name_bytes: bytes) -> int:
        return int(self.ctx.lib.mmdgpu_model_find_bone(self.h, name_bytes, len(name_bytes)))

    def find_morph(self, name_bytes: bytes) -> int:
        return int(self.ctx.lib.mmdgpu_model_find_morph(self.h, name_bytes, len(name_bytes)))

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.mmdgpu_model_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Motion:
    """Flattened VMD clip bound to one model (mmd::Motion + MotionPlayer's name join)."""

    def __init__(self, model: Model, arrays: dict | None = None, vmd_bytes: bytes | None = None):
        self.model = model
        ctx = model.ctx
        h = C.c_void_p()
        if arrays is not None:
            desc, keep = capi.anim_desc(arrays)
            check(ctx.lib.mmdgpu_animation_create_from_arrays(ctx.h, model.h, C.byref(desc), C.byref(h)), ctx.h)
            del keep
        elif vmd_bytes is not None:
            buf = (C.c_char * len(vmd_bytes)).from_buffer_copy(vmd_bytes)
            check(ctx.lib.mmdgpu_animation_create_from_vmd(ctx.h, model.h, buf, len(vmd_bytes), C.byref(h)), ctx.h)
        else:
            raise ValueError("arrays or vmd_bytes required")
        self.h = h

    def GetLength(self) -> int:
        return int(self.model.ctx.lib.mmdgpu_animation_length(self.h))

    def close(self):
        if getattr(self, "h", None) and self.model.ctx.h:
            self.model.ctx.lib.mmdgpu_animation_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Frames:
    """n_instances x n_frames independent Poser states evaluated per launch (mmdgpu_frames_t)."""

    def __init__(self, model: Model, n_instances: int = 1, n_frames: int = 1,
                 layout: int = capi.LAYOUT_SOA_POS_NRM):
        self.model = model
        self.ctx = model.ctx
        self.lib = self.ctx.lib
        self.layout = layout
        self.n_instances, self.n_frames = int(n_instances), int(n_frames)
        self.n_slots = self.n_instances * self.n_frames
        h = C.c_void_p()
        check(self.lib.mmdgpu_frames_create(self.ctx.h, model.h, self.n_instances, self.n_frames, int(layout),
                                            C.byref(h)), self.ctx.h)
        self.h = h

    # ---- libmmd-named entry points, batched
    def anim_array(self, motions):
        """Marshal one clip per instance once; the result can be passed wherever `motions` is accepted (a crowd step that
        re-marshals 512 handles per call spends more time in ctypes than on the device)."""
        return self._anim_array(motions)

    def _anim_array(self, motions):
        if isinstance(motions, C.Array):
            if len(motions) != self.n_instances:
                raise ValueError("one motion per instance required")
            return motions
        if isinstance(motions, Motion):
            motions = [motions] * self.n_instances
        if len(motions) != self.n_instances:
            raise ValueError("one motion per instance required")
        arr = (C.c_void_p * self.n_instances)(*[m.h for m in motions])
        return arr

    def reset_posing(self):
        check(self.lib.mmdgpu_reset_posing(self.h), self.ctx.h)

    def seek_frame(self, motions, frame_per_slot):
        f = np.ascontiguousarray(frame_per_slot, np.uint32)
        if f.size != self.n_slots:
            raise ValueError("one frame id per slot required")
        check(self.lib.mmdgpu_seek_frame(self.h, self._anim_array(motions), _ptr(f)), self.ctx.h)

    def seek_frame_range(self, motions, first_frame_per_instance, stride: int = 1):
        f = np.ascontiguousarray(first_frame_per_instance, np.uint32)
        if f.size != self.n_instances:
            raise ValueError("one first frame per instance required")
        check(self.lib.mmdgpu_seek_frame_range(self.h, self._anim_array(motions), _ptr(f), int(stride)), self.ctx.h)

    def seek_time(self, motions, time_per_slot):
        t = np.ascontiguousarray(time_per_slot, np.float64)
        if t.size != self.n_slots:
            raise ValueError("one time per slot required")
        check(self.lib.mmdgpu_seek_time(self.h, self._anim_array(motions), _ptr(t)), self.ctx.h)

    def pose_frame(self, motions, frame_per_slot):
        """reset_posing + seek_frame + pre_physics_posing + post_physics_posing in one call (one sampling launch, one hierarchy pass)."""
        f = np.ascontiguousarray(frame_per_slot, np.uint32)
        if f.size != self.n_slots:
            raise ValueError("one frame id per slot required")
        check(self.lib.mmdgpu_pose_frame(self.h, self._anim_array(motions), _ptr(f)), self.ctx.h)

    def pose_time(self, motions, time_per_slot):
        t = np.ascontiguousarray(time_per_slot, np.float64)
        if t.size != self.n_slots:
            raise ValueError("one time per slot required")
        check(self.lib.mmdgpu_pose_time(self.h, self._anim_array(motions), _ptr(t)), self.ctx.h)

    def reset_and_seek_frame(self, motions, frame_per_slot):
        """reset_posing + seek_frame as one sampling launch (unanimated bones / morphs get identity / zero)."""
        f = np.ascontiguousarray(frame_per_slot, np.uint32)
        if f.size != self.n_slots:
            raise ValueError("one frame id per slot required")
        check(self.lib.mmdgpu_reset_and_seek_frame(self.h, self._anim_array(motions), _ptr(f)), self.ctx.h)

    def reset_and_seek_time(self, motions, time_per_slot):
        t = np.ascontiguousarray(time_per_slot, np.float64)
        if t.size != self.n_slots:
            raise ValueError("one time per slot required")
        check(self.lib.mmdgpu_reset_and_seek_time(self.h, self._anim_array(motions), _ptr(t)), self.ctx.h)

    def set_bone_pose(self, slot: int, bone: int, translation, rotation):
        t = np.ascontiguousarray(translation, np.float32)
        r = np.ascontiguousarray(rotation, np.float32)
        check(self.lib.mmdgpu_set_bone_pose(self.h, int(slot), int(bone), _ptr(t), _ptr(r)), self.ctx.h)

    def set_morph_pose(self, slot: int, morph: int, weight: float):
        check(self.lib.mmdgpu_set_morph_pose(self.h, int(slot), int(morph), float(weight)), self.ctx.h)

    def pre_physics_posing(self):
        check(self.lib.mmdgpu_pre_physics_posing(self.h), self.ctx.h)

    def post_physics_posing(self):
        check(self.lib.mmdgpu_post_physics_posing(self.h), self.ctx.h)

    def deform(self):
        check(self.lib.mmdgpu_deform(self.h), self.ctx.h)

    def set_skinning_matrix_override(self, slot: int, bone: int, skinning16, local16=None):
        s = np.ascontiguousarray(skinning16, np.float32).reshape(16)
        l = None if local16 is None else np.ascontiguousarray(local16, np.float32).reshape(16)
        check(self.lib.mmdgpu_set_skinning_matrix_override(self.h, int(slot), int(bone), _ptr(s), _ptr(l)), self.ctx.h)

    def update(self, motions, frame_per_slot):
        f = np.ascontiguousarray(frame_per_slot, np.uint32)
        if f.size != self.n_slots:
            raise ValueError("one frame id per slot required")
        check(self.lib.mmdgpu_update(self.h, self._anim_array(motions), _ptr(f)), self.ctx.h)

    def update_range(self, motions, first_frame_per_instance, stride: int = 1):
        f = np.ascontiguousarray(first_frame_per_instance, np.uint32)
        if f.size != self.n_instances:
            raise ValueError("one first frame per instance required")
        check(self.lib.mmdgpu_update_range(self.h, self._anim_array(motions), _ptr(f), int(stride)), self.ctx.h)

    # ---- outputs
    def device_ptr(self, stream_id: int):
        p, stride = C.c_void_p(), C.c_size_t()
        check(self.lib.mmdgpu_frames_device_ptr(self.h, int(stream_id), C.byref(p), C.byref(stride)), self.ctx.h)
        return int(p.value or 0), int(stride.value)

    def download(self, slot: int, stream_id: int) -> np.ndarray:
        nv, nb = self.model.n_vertices, self.model.n_bones
        shape = {capi.STREAM_POSITION: (nv, 3), capi.STREAM_NORMAL: (nv, 3), capi.STREAM_INTERLEAVED: (nv, 8),
                 capi.STREAM_SKIN_MATRIX: (nb, 12), capi.STREAM_UV: (nv, 2)}[stream_id]
        out = np.empty(shape, np.float32)
        check(self.lib.mmdgpu_frames_download(self.h, int(slot), int(stream_id), _ptr(out), out.nbytes), self.ctx.h)
        return out

    def download_async(self, first_slot: int, n_slots: int, stream_id: int, pinned_ptr: int, nbytes: int):
        check(self.lib.mmdgpu_frames_download_async(self.h, int(first_slot), int(n_slots), int(stream_id),
                                                    C.c_void_p(pinned_ptr), int(nbytes)), self.ctx.h)

    @property
    def slots_per_cta(self) -> int:
        return int(self.lib.mmdgpu_frames_slot_run(self.h))

    def download_pair_async(self, slot: int, pinned_ptr: int, nbytes: int):
        check(self.lib.mmdgpu_frames_download_pair_async(self.h, int(slot), C.c_void_p(pinned_ptr), int(nbytes)), self.ctx.h)

    def wait_skinning(self):
        """Blocks until the latest skinning launch of this object has finished (bound host outputs are then complete)."""
        check(self.lib.mmdgpu_frames_wait_skinning(self.h), self.ctx.h)

    def wait_downloads(self):
        """Host-blocks until this object's download_async copies have landed (not compute, not other objects)."""
        check(self.lib.mmdgpu_frames_wait_downloads(self.h), self.ctx.h)

    def downloads_done(self) -> bool:
        r = int(self.lib.mmdgpu_frames_downloads_done(self.h))
        if r < 0:
            raise MmdGpuError(capi.ERR_CUDA, "cudaEventQuery failed")
        return r == 1

    def bind_output(self, stream_id: int, device_ptr: int | None, slot_stride_bytes: int = 0):
        """Let the skinning kernel write `stream_id` into caller-owned device memory (None: library buffer again)."""
        check(self.lib.mmdgpu_frames_bind_output(self.h, int(stream_id), C.c_void_p(device_ptr) if device_ptr else None,
                                                 int(slot_stride_bytes)), self.ctx.h)

    def bone_matrices(self, slot: int = 0) -> np.ndarray:
        out = np.empty((self.model.n_bones, 16), np.float32)
        check(self.lib.mmdgpu_bone_matrices_download(self.h, int(slot), _ptr(out)), self.ctx.h)
        return out

    def bone_local_matrices(self, slot: int = 0) -> np.ndarray:
        out = np.empty((self.model.n_bones, 16), np.float32)
        check(self.lib.mmdgpu_bone_local_matrices_download(self.h, int(slot), _ptr(out)), self.ctx.h)
        return out

    def bone_poses(self, slot: int = 0) -> np.ndarray:
        out = np.empty((self.model.n_bones, 7), np.float32)
        check(self.lib.mmdgpu_bone_poses_download(self.h, int(slot), _ptr(out)), self.ctx.h)
        return out

    def morph_rates(self, slot: int = 0) -> np.ndarray:
        out = np.empty((self.model.n_morphs,), np.float32)
        check(self.lib.mmdgpu_morph_rates_download(self.h, int(slot), _ptr(out)), self.ctx.h)
        return out

    def material_images(self, slot: int = 0) -> np.ndarray:
        """Poser::material_mul_images_ / material_add_images_ (poser.inl:107-161): (n_materials, 2, 28); [:, 0] is the
        multiplicative image, [:, 1] the additive one.  All 1 / all 0 in libmmd-exact mode (libmmd never fills them)."""
        out = np.empty((self.model.n_materials, 2, capi.MATERIAL_FIELDS), np.float32)
        check(self.lib.mmdgpu_material_images_download(self.h, int(slot), _ptr(out)), self.ctx.h)
        return out

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.lib.mmdgpu_frames_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PoseImage:
    """Poser::pose_image (L/motion/poser.inl:17-20): lazily downloaded host mirror of the deformed buffer."""

    def __init__(self, frames: Frames):
        self._f = frames
        self._coords = self._normals = None

    def invalidate(self):
        self._coords = self._normals = None

    @property
    def coordinates(self) -> np.ndarray:
        if self._coords is None:
            self._coords = self._f.download(0, capi.STREAM_POSITION)
        return self._coords

    @property
    def normals(self) -> np.ndarray:
        if self._normals is None:
            self._normals = self._f.download(0, capi.STREAM_NORMAL)
        return self._normals


class Poser:
    """mmd::Poser with libmmd's method names (one slot on the device)."""

    def __init__(self, model: Model):
        self.model = model
        self.frames = Frames(model, 1, 1, capi.LAYOUT_SOA_POS_NRM)
        self.pose_image = PoseImage(self.frames)
        # Poser::Poser ends with ResetPosing(); Deform(); (poser_impl.inl:125-127)
        self.ResetPosing()
        self.Deform()

    def GetModel(self) -> Model:
        return self.model

    def ResetPosing(self):
        """poser_impl.inl:130-140, including the Pre+PostPhysicsPosing evaluation it ends with."""
        self.frames.reset_posing()
        self.frames.pre_physics_posing()
        self.frames.post_physics_posing()

    def SetBonePose(self, index: int, translation, rotation):
        self.frames.set_bone_pose(0, index, translation, rotation)

    def SetMorphPose(self, index: int, weight: float):
        self.frames.set_morph_pose(0, index, weight)

    def PrePhysicsPosing(self):
        self.frames.pre_physics_posing()

    def PostPhysicsPosing(self):
        self.frames.post_physics_posing()

    def Deform(self):
        self.frames.deform()
        self.pose_image.invalidate()

    def skinning_matrices(self) -> np.ndarray:
        return self.frames.bone_matrices(0)


class MotionPlayer:
    """mmd::MotionPlayer (L/motion/poser.inl:184-198, poser_impl.inl:522-546)."""

    def __init__(self, motion: Motion, poser: Poser):
        if motion.model is not poser.model:
            raise ValueError("motion and poser were built for different models")
        self.motion, self.poser = motion, poser

    def SeekFrame(self, frame: int):
        self.poser.frames.seek_frame(self.motion, [int(frame)])

    def SeekTime(self, seconds: float):
        """MotionPlayer::SeekTime(double), poser_impl.inl:548-555."""
        self.poser.frames.seek_time(self.motion, [float(seconds)])


# ------------------------------------------------------------------------------------ host-only plan access
def _plan_to_dict(lib, plan_h) -> dict:
    out = {}
    for which, dt in capi.PLAN_DTYPES.items():
        p, n = C.c_void_p(), C.c_size_t()
        check(lib.mmdgpu_plan_get(plan_h, which, C.byref(p), C.byref(n)))
        cnt = int(n.value)
        if cnt == 0:
            out[which] = np.zeros(0, dt)
        else:
            buf = (C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(p.value)
            out[which] = np.frombuffer(buf, dtype=dt).copy()
    return out


def plan_arrays(model_arrays: dict, extensions: bool = False) -> dict:
    """Host-only flattening of a model (no GPU): dict keyed by capi.PLAN_* (index-tier parity checks)."""
    lib = load()
    desc, keep = capi.model_desc(model_arrays)
    opt = capi.Options()
    opt.extensions = 1 if extensions else 0
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    st = lib.mmdgpu_plan_create(C.byref(desc), C.byref(opt), C.byref(h), err, 512)
    if st != capi.OK:
        raise MmdGpuError(st, err.value.decode("utf-8", "replace"))
    try:
        return _plan_to_dict(lib, h)
    finally:
        lib.mmdgpu_plan_destroy(h)
        del keep


def _arrays_from(getter, handle, dtypes) -> dict:
    out = {}
    for which, dt in dtypes.items():
        p, n = C.c_void_p(), C.c_size_t()
        check(getter(handle, which, C.byref(p), C.byref(n)))
        cnt = int(n.value)
        if cnt == 0:
            out[which] = np.zeros(0, dt)
        else:
            buf = (C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(p.value)
            out[which] = np.frombuffer(buf, dtype=dt).copy()
    return out


class HostPlan:
    """Host-only plan handle (no GPU): from flat arrays or PMX bytes; `anim_*` flatten motions against it."""

    def __init__(self, arrays: dict | None = None, pmx_bytes: bytes | None = None, extensions: bool = False):
        self.lib = load()
        opt = capi.Options()
        opt.extensions = 1 if extensions else 0
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        if arrays is not None:
            desc, keep = capi.model_desc(arrays)
            st = self.lib.mmdgpu_plan_create(C.byref(desc), C.byref(opt), C.byref(h), err, 512)
        else:
            buf = (C.c_char * len(pmx_bytes)).from_buffer_copy(pmx_bytes)
            st = self.lib.mmdgpu_plan_create_from_pmx(buf, len(pmx_bytes), C.byref(opt), C.byref(h), err, 512)
        if st != capi.OK:
            raise MmdGpuError(st, err.value.decode("utf-8", "replace"))
        self.h = h

    def arrays(self) -> dict:
        return _plan_to_dict(self.lib, self.h)

    def anim_from_arrays(self, motion: dict, n_bones: int, n_morphs: int) -> dict:
        desc, keep = capi.anim_desc(motion)
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        st = self.lib.mmdgpu_anim_plan_create(C.byref(desc), int(n_bones), int(n_morphs), C.byref(h), err, 512)
        if st != capi.OK:
            raise MmdGpuError(st, err.value.decode("utf-8", "replace"))
        try:
            return _arrays_from(self.lib.mmdgpu_anim_plan_get, h, capi.ANIM_DTYPES)
        finally:
            self.lib.mmdgpu_anim_plan_destroy(h)

    def anim_from_vmd(self, vmd_bytes: bytes) -> dict:
        buf = (C.c_char * len(vmd_bytes)).from_buffer_copy(vmd_bytes)
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        st = self.lib.mmdgpu_anim_plan_create_from_vmd(self.h, buf, len(vmd_bytes), C.byref(h), err, 512)
        if st != capi.OK:
            raise MmdGpuError(st, err.value.decode("utf-8", "replace"))
        try:
            return _arrays_from(self.lib.mmdgpu_anim_plan_get, h, capi.ANIM_DTYPES)
        finally:
            self.lib.mmdgpu_anim_plan_destroy(h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.mmdgpu_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bezier_table(ctrl4) -> np.ndarray | None:
    """Bezier::presample table for one VMD control quadruple; None if the curve is linear."""
    lib = load()
    c = np.ascontiguousarray(ctrl4, np.int8)
    t = np.zeros(32, np.float32)
    return None if lib.mmdgpu_bezier_table(_ptr(c), _ptr(t)) else t
