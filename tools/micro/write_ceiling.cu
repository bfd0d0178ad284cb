// Micro-benchmark (sm_100a): how fast can a write-dominated kernel push bytes to HBM on this GPU?  The skinning kernel's
// DRAM traffic is 96 % writes (24 B per vertex-frame out, static streams amortised over a 64-slot run), so its HBM
// roofline is the WRITE ceiling of its own output mechanism, not the read+write copy peak.  Variants:
//   memset        cudaMemsetAsync
//   stg128        grid-stride st.global.v4.f32 (plain coalesced stores)
//   bulk6k        the skinning kernel's output path with the compute removed: CTAs of 128 threads, 3 per SM, each owning a
//                 512-vertex tile and walking `slots` slots; per slot two cp.async.bulk shared->global copies of 6 KB
//                 (position plane, normal plane) at the same addresses the real kernel writes, 4 slots per commit group
//   stg256        the sokol32 path: one st.global.v8.f32 (32-byte record) per lane at scattered positions inside the tile
//   copy          cudaMemcpyAsync device->device (the MEASURED_PEAKS.json method), for reference
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_ceiling write_ceiling.cu && ./write_ceiling
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void stg128(float4* out, size_t n4) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr uint32_t kTile = 512, kG = 4;

__global__ void __launch_bounds__(128, 3) bulk6k(float* out_pos, float* out_nrm, uint32_t n_tiles, uint32_t slots, uint32_t chunk,
                                                 size_t slot_stride_floats) {
    extern __shared__ __align__(128) unsigned char sm[];   // kG staging tiles of 12 KB
    const uint32_t n_chunks = (slots + chunk - 1) / chunk;
    const uint32_t tile = blockIdx.x / n_chunks, ck = blockIdx.x % n_chunks;
    if (tile >= n_tiles) return;
    const uint32_t s0 = ck * chunk, s1 = min(slots, s0 + chunk);
    for (uint32_t i = threadIdx.x; i < kG * kTile * 6; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (float)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (uint32_t g0 = s0; g0 < s1; g0 += kG) {
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            for (uint32_t f = 0; f < kG && g0 + f < s1; ++f) {
                unsigned char* stage = sm + (size_t)f * kTile * 24;
                float* dp = out_pos + (size_t)(g0 + f) * slot_stride_floats + (size_t)tile * kTile * 3;
                float* dn = out_nrm + (size_t)(g0 + f) * slot_stride_floats + (size_t)tile * kTile * 3;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dp), "r"(smem_u32(stage)), "r"(kTile * 12u) : "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dn), "r"(smem_u32(stage + kTile * 12)), "r"(kTile * 12u) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(128, 3) stg256(float4* out, uint32_t n_tiles, uint32_t slots, uint32_t chunk, size_t slot_stride_f4) {
    const uint32_t n_chunks = (slots + chunk - 1) / chunk;
    const uint32_t tile = blockIdx.x / n_chunks, ck = blockIdx.x % n_chunks;
    if (tile >= n_tiles) return;
    const uint32_t s0 = ck * chunk, s1 = min(slots, s0 + chunk);
    for (uint32_t s = s0; s < s1; ++s)
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            // a class-sorted warp step writes an irregular subset of the tile: here a fixed pseudo-random permutation
            const uint32_t o = ((threadIdx.x * 4 + j) * 197u + 31u) & (kTile - 1);
            float4* p = out + (size_t)s * slot_stride_f4 + ((size_t)tile * kTile + o) * 2u;
            asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f), "f"(5.f),
                         "f"(6.f), "f"(7.f), "f"(8.f) : "memory");
        }
}

int main() {
    const uint32_t nv = 1000448, n_tiles = nv / kTile, slots = 128, chunk = 64;   // the C3 headline launch
    const size_t plane = (size_t)slots * nv * 12;                                   // 1.54 GB per plane
    float *pos, *nrm;
    CK(cudaMalloc(&pos, plane));
    CK(cudaMalloc(&nrm, plane));
    float4* inter;
    CK(cudaMalloc(&inter, (size_t)slots * nv * 32));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(bulk6k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    auto report = [&](const char* name, double bytes, float ms, double rw = 1.0) {
        printf("%-44s %8.3f ms  %8.1f GB/s written%s\n", name, ms, bytes / (ms * 1e-3) / 1e9, rw > 1.0 ? " (+ as many read)" : "");
    };
    const int reps = 5;
    float ms;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) { CK(cudaMemsetAsync(pos, 1, plane)); CK(cudaMemsetAsync(nrm, 2, plane)); }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("memset (2 x 1.54 GB)", 2.0 * plane, ms / reps);

        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) { stg128<<<148 * 16, 256>>>((float4*)pos, plane / 16); stg128<<<148 * 16, 256>>>((float4*)nrm, plane / 16); }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("stg128 grid-stride (2 x 1.54 GB)", 2.0 * plane, ms / reps);

        for (uint32_t ch : {64u, 16u, 128u}) {
            const uint32_t n_chunks = (slots + ch - 1) / ch;
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) bulk6k<<<n_tiles * n_chunks, 128, kG * kTile * 24>>>(pos, nrm, n_tiles, slots, ch, (size_t)nv * 3);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            char name[96];
            snprintf(name, sizeof name, "bulk 6 KB tiles, %u-slot runs (3.07 GB)", ch);
            if (pass) report(name, 2.0 * plane, ms / reps);
        }
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) stg256<<<n_tiles * 2, 128>>>(inter, n_tiles, slots, chunk, (size_t)nv * 2);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("stg256 scattered 32-B records (4.10 GB)", (double)slots * nv * 32, ms / reps);

        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(nrm, pos, plane, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("memcpy d2d (1.54 GB)", (double)plane, ms / reps, 2.0);
    }
    CK(cudaGetLastError());
    return 0;
}
