// Micro-benchmark (sm_100a): how many shared-memory wavefronts does one LDS.128 / LDS.32 / STS.32 cost for the lane
// address patterns the skinning kernel produces?  Each pattern is a table of 32 per-lane element indices; the kernel
// issues `iters` x 4 independent accesses per warp with 16 warps per SM, so the LSU data pipe is the limiter and
// cycles / instruction ~ wavefronts / instruction.  Run under ncu for the exact wavefront counters.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_patterns smem_patterns.cu && ./smem_patterns
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int kWarps = 16, kIters = 2048;

__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ float2 lds64(const float2* p) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ float lds32(const float* p) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ void sts32(float* p, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v) : "memory");
}

template <int MODE>  // 0 LDS.128, 1 LDS.32, 2 STS.32 x3 (12-byte records), 3 LDS.64
__global__ void __launch_bounds__(kWarps * 32) k(const int* __restrict__ table, float* out, long long* cycles) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int idx = table[lane];
    float acc = 0.f;
    const long long t0 = clock64();
    if (MODE == 0) {
        const float4* p = reinterpret_cast<const float4*>(sm) + idx;
#pragma unroll 1
        for (int i = 0; i < kIters; ++i) {
            // + multiples of 32 float4 (512 B) keep the bank pattern; the xor keeps the loads from being hoisted
            const int o = (i & 3) * 32;
            const float4 a = lds128(p + o), b = lds128(p + o + 128), c = lds128(p + o + 256), d = lds128(p + o + 384);
            acc += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (d.x + d.y + d.z + d.w);
        }
    } else if (MODE == 1) {
        const float* p = sm + idx;
#pragma unroll 1
        for (int i = 0; i < kIters; ++i) {
            const int o = (i & 3) * 32;
            acc += lds32(p + o) + lds32(p + o + 512) + lds32(p + o + 1024) + lds32(p + o + 1536);
        }
    } else if (MODE == 3) {
        const float2* p = reinterpret_cast<const float2*>(sm) + idx;
#pragma unroll 1
        for (int i = 0; i < kIters; ++i) {
            const int o = (i & 3) * 32;
            const float2 a = lds64(p + o), b = lds64(p + o + 256), c = lds64(p + o + 512), d = lds64(p + o + 768);
            acc += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
        }
    } else {
        float* p = sm + 3 * idx;
#pragma unroll 1
        for (int i = 0; i < kIters; ++i) {
            const float v = (float)i;
            sts32(p, v); sts32(p + 1, v); sts32(p + 2, v);
            sts32(p + 3 * 512, v); sts32(p + 3 * 512 + 1, v); sts32(p + 3 * 512 + 2, v);  // second tile: same banks (1536 words = 48 x 32)
        }
        __syncthreads();
        acc = sm[threadIdx.x];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Pattern { const char* name; int mode; int idx[32]; };

int main() {
    std::vector<Pattern> pats;
    auto add = [&](const char* name, int mode, auto f) {
        Pattern p{name, mode, {}};
        for (int l = 0; l < 32; ++l) p.idx[l] = f(l);
        pats.push_back(p);
    };
    srand(12345);
    int rnd12[32], rnd12s[32], rnd7[32], rnd7s[32], rnd24[32], rnd24s[32];
    for (int l = 0; l < 32; ++l) { rnd12[l] = rand() % 12; rnd7[l] = rand() % 7; rnd24[l] = rand() % 24; }
    std::copy(rnd12, rnd12 + 32, rnd12s); std::sort(rnd12s, rnd12s + 32);
    std::copy(rnd7, rnd7 + 32, rnd7s); std::sort(rnd7s, rnd7s + 32);
    std::copy(rnd24, rnd24 + 32, rnd24s); std::sort(rnd24s, rnd24s + 32);
    // ---- LDS.128, index in float4 units
    add("lds128 uniform (all lanes one address)", 0, [](int) { return 0; });
    add("lds128 consecutive (32 distinct)", 0, [](int l) { return l; });
    add("lds128 lane/4 (8 distinct, runs of 4)", 0, [](int l) { return l / 4; });
    add("lds128 lane%8 (8 distinct, interleaved)", 0, [](int l) { return l % 8; });
    add("lds128 lane/8 (quarter-warps uniform, 4 distinct)", 0, [](int l) { return l / 8; });
    add("lds128 lane/16 (2 distinct)", 0, [](int l) { return l / 16; });
    add("lds128 lane%2 (2 distinct interleaved)", 0, [](int l) { return l % 2; });
    add("lds128 lane/2 (16 distinct pairs)", 0, [](int l) { return l / 2; });
    add("lds128 palette stride 3: bone = lane%8", 0, [](int l) { return 3 * (l % 8); });
    add("lds128 palette stride 3: bone = lane/4", 0, [](int l) { return 3 * (l / 4); });
    add("lds128 palette stride 3: bone = lane/16", 0, [](int l) { return 3 * (l / 16); });
    add("lds128 palette stride 3: 12 bones random", 0, [&](int l) { return 3 * rnd12[l]; });
    add("lds128 palette stride 3: 12 bones sorted", 0, [&](int l) { return 3 * rnd12s[l]; });
    add("lds128 palette stride 3: 7 bones random", 0, [&](int l) { return 3 * rnd7[l]; });
    add("lds128 palette stride 3: 7 bones sorted", 0, [&](int l) { return 3 * rnd7s[l]; });
    add("lds128 palette stride 3: 24 bones random", 0, [&](int l) { return 3 * rnd24[l]; });
    add("lds128 palette stride 3: 24 bones sorted", 0, [&](int l) { return 3 * rnd24s[l]; });
    add("lds128 palette stride 3: 32 distinct bones", 0, [](int l) { return 3 * l; });
    // is the pair fast path all-or-nothing, or per quarter-warp?
    add("lds128 q0 aligned pairs, rest distinct", 0, [](int l) { return l < 8 ? l / 2 : l; });
    add("lds128 q0-q1 aligned pairs, rest distinct", 0, [](int l) { return l < 16 ? l / 2 : l; });
    add("lds128 aligned pairs except one mixed pair", 0, [](int l) { return l == 31 ? 40 : l / 2; });
    add("lds128 pairs shifted by one lane ((l+1)/2)", 0, [](int l) { return (l + 1) / 2; });
    add("lds128 stride 3: sorted runs, boundaries on even lanes", 0, [](int l) { const int b[16] = {0,0,1,1,1,2,3,3,3,3,4,5,5,6,6,7}; return 3 * b[l / 2]; });
    // ---- LDS.64, index in float2 units
    add("lds64 consecutive", 3, [](int l) { return l; });
    add("lds64 12 bones random, stride 6", 3, [&](int l) { return 6 * rnd12[l]; });
    add("lds64 12 bones sorted, stride 6", 3, [&](int l) { return 6 * rnd12s[l]; });
    add("lds64 uniform", 3, [](int) { return 0; });
    // planar palette for slot PAIRS: one element of bone b for two slots = one 8-byte cell at float2 index b
    add("lds64 planar: 7 bones random", 3, [&](int l) { return rnd7[l]; });
    add("lds64 planar: 12 bones random", 3, [&](int l) { return rnd12[l]; });
    add("lds64 planar: 12 bones sorted", 3, [&](int l) { return rnd12s[l]; });
    add("lds64 planar: 16 bones (lane % 16)", 3, [](int l) { return l % 16; });
    add("lds64 planar: 24 bones random", 3, [&](int l) { return rnd24[l]; });
    add("lds64 planar: lane / 2", 3, [](int l) { return l / 2; });
    // the same for LDS.128 planar (two elements of a bone for two slots in one 16-byte cell at float4 index b)
    add("lds128 planar: 7 bones random", 0, [&](int l) { return rnd7[l]; });
    add("lds128 planar: 12 bones random", 0, [&](int l) { return rnd12[l]; });
    add("lds128 planar: lane % 4", 0, [](int l) { return l % 4; });
    add("lds128 planar: 4 bones random", 0, [&](int l) { return rnd12[l] % 4; });
    // ---- LDS.32, index in words (planar palette: one element of bone b at word b)
    add("lds32 consecutive", 1, [](int l) { return l; });
    add("lds32 planar: 12 bones random", 1, [&](int l) { return rnd12[l]; });
    add("lds32 planar: 24 bones random", 1, [&](int l) { return rnd24[l]; });
    add("lds32 stride 12 words: 12 bones random", 1, [&](int l) { return 12 * rnd12[l]; });
    // ---- STS.32 x 3 at 12-byte records, index = record
    add("sts32x3 records consecutive", 2, [](int l) { return l; });
    {
        // an increasing irregular subset of a 512-record tile at density 1/3 (what a class-sorted warp writes)
        int sub[32], o = 0;
        for (int l = 0; l < 32; ++l) { o += 1 + rand() % 5; sub[l] = o % 512; }
        add("sts32x3 records irregular subset (density 1/3)", 2, [&](int l) { return sub[l]; });
        // the same subset after a residue-dealt reorder is 'distinct mod 32'
        add("sts32x3 records distinct mod 32, scattered", 2, [](int l) { return (l * 37 + (l % 5) * 32) % 512; });
    }
    int* d_table; float* d_out; long long* d_cyc;
    cudaMalloc(&d_table, 32 * sizeof(int));
    cudaMalloc(&d_out, 148 * kWarps * 32 * sizeof(float));
    cudaMalloc(&d_cyc, 148 * sizeof(long long));
    for (auto& p : pats) {
        cudaMemcpy(d_table, p.idx, sizeof(p.idx), cudaMemcpyHostToDevice);
        for (int rep = 0; rep < 2; ++rep) {
            const size_t sh = 8192 * 4;
            if (p.mode == 0) k<0><<<148, kWarps * 32, sh>>>(d_table, d_out, d_cyc);
            else if (p.mode == 1) k<1><<<148, kWarps * 32, sh>>>(d_table, d_out, d_cyc);
            else if (p.mode == 3) k<3><<<148, kWarps * 32, sh>>>(d_table, d_out, d_cyc);
            else k<2><<<148, kWarps * 32, sh>>>(d_table, d_out, d_cyc);
        }
        cudaDeviceSynchronize();
        long long cyc[148];
        cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (long long c : cyc) mean += (double)c / 148;
        const double instr = (double)kIters * (p.mode == 2 ? 6 : 4) * kWarps;
        printf("%-58s %7.3f cycles / warp instruction\n", p.name, mean / instr);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
