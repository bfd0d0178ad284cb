// kernels.cuh — host-callable launchers of the sm_100a kernels (kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include "device_types.cuh"

namespace mmdgpu {

// What to sample: anims == nullptr writes identity / zero (Poser::ResetPosing's pose part); write_untracked: also write
// identity / zero for items the clip does not animate; frame_by_value / time_by_value: one frame id (first frame of a
// one-instance range) or one time handed over as a kernel argument instead of through F.frame_id / F.time_s;
// frame_ids_inline / n_inline: up to kInlineFrameIds frame ids (per slot, or first frames per instance in range mode) handed
// over the same way - a small crowd's step is tens of microseconds of device work, and the host-to-device copy of its
// ids from pageable memory is the most expensive driver call of the update.
constexpr uint32_t kInlineFrameIds = 64;
struct SampleSpec {
    const DevAnim* anims = nullptr;
    bool write_untracked = false, range_mode = false;
    uint32_t frame_stride = 1;
    bool time_mode = false;
    const uint32_t* frame_by_value = nullptr;
    const double* time_by_value = nullptr;
    const uint32_t* frame_ids_inline = nullptr;
    uint32_t n_inline = 0;
};
// K1: keyframe sampling for every (slot, bone) and (slot, morph).
cudaError_t launch_pose_sample(cudaStream_t st, const DevModel& M, const DevFrames& F, const SampleSpec& spec);
// K2: waves [wave_lo, wave_hi) of the bone program; prologue = morph rates + reset + bone morphs.
cudaError_t launch_hierarchy(cudaStream_t st, const DevModel& M, const DevFrames& F, uint32_t wave_lo, uint32_t wave_hi,
                             bool prologue);
// K2, one wave with a thread per (op, slot): CCD IK solves on chain-local images (kernels.cu), for large batches.
bool hierarchy_uses_cta_kernel(const DevModel& M);
size_t hierarchy_flat_smem_bytes(const DevModel& M);
cudaError_t launch_hierarchy_wave_flat(cudaStream_t st, const DevModel& M, const DevFrames& F, uint32_t wave, uint32_t n_ops);
// K3: skinning + fused vertex-morph gather.  One CTA = one 512-vertex tile x `slots_per_cta` consecutive slots, four at a time.
size_t skin_smem_bytes(const DevModel& M, int layout);
cudaError_t prepare_skin_kernels(const DevModel& M);
cudaError_t launch_skin(cudaStream_t st, const DevModel& M, const DevFrames& F, int layout, uint32_t slots_per_cta);

// Test export: one device math function per case (kernels.cu, math_kat_kernel).
cudaError_t launch_math_kat(cudaStream_t st, int op, const float* in, uint32_t n, float* out, const float* tables, const uint32_t* curve);

}  // namespace mmdgpu
