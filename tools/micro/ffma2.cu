// Micro-benchmark: issue rate of packed fp32 (fma.rn.f32x2 -> FFMA2) vs scalar FMUL+FADD on sm_100a, and a
// bit-exactness check of  mul = fma(a, b, -0)  and  add = fma(a, 1, b)  against un-fused scalar ops.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    if (MODE == 0) {  // scalar: 8 independent chains of FMUL then FADD (un-fused)
        for (int i = 0; i < iters; ++i) {
            a0 = __fadd_rn(__fmul_rn(a0, s), 1.0f); a1 = __fadd_rn(__fmul_rn(a1, s), 1.0f);
            a2 = __fadd_rn(__fmul_rn(a2, s), 1.0f); a3 = __fadd_rn(__fmul_rn(a3, s), 1.0f);
            a4 = __fadd_rn(__fmul_rn(a4, s), 1.0f); a5 = __fadd_rn(__fmul_rn(a5, s), 1.0f);
            a6 = __fadd_rn(__fmul_rn(a6, s), 1.0f); a7 = __fadd_rn(__fmul_rn(a7, s), 1.0f);
        }
    } else {          // packed: 4 chains of (mul as fma(a,s,-0)) then (add as fma(x,1,1)) on float2 pairs
        unsigned long long p0 = pk(a0, a1), p1 = pk(a2, a3), p2 = pk(a4, a5), p3 = pk(a6, a7);
        const unsigned long long S = pk(s, s), NZ = pk(-0.0f, -0.0f), ONE = pk(1.0f, 1.0f);
        for (int i = 0; i < iters; ++i) {
            p0 = fma2(fma2(p0, S, NZ), ONE, ONE); p1 = fma2(fma2(p1, S, NZ), ONE, ONE);
            p2 = fma2(fma2(p2, S, NZ), ONE, ONE); p3 = fma2(fma2(p3, S, NZ), ONE, ONE);
        }
        upk(p0, a0, a1); upk(p1, a2, a3); upk(p2, a4, a5); upk(p3, a6, a7);
    }
    out[(blockIdx.x * blockDim.x + threadIdx.x)] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void exact(const float* a, const float* b, const float* c, uint32_t* mism, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = a[i], y = b[i], z = c[i];
    float m = __fmul_rn(x, y), s = __fadd_rn(m, z);
    float pm0, pm1, ps0, ps1;
    upk(fma2(pk(x, y), pk(y, x), pk(-0.0f, -0.0f)), pm0, pm1);
    upk(fma2(pk(m, z), pk(1.0f, 1.0f), pk(z, m)), ps0, ps1);
    if (__float_as_uint(pm0) != __float_as_uint(m) || __float_as_uint(pm1) != __float_as_uint(m) ||
        __float_as_uint(ps0) != __float_as_uint(s) || __float_as_uint(ps1) != __float_as_uint(s)) atomicAdd(mism, 1u);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * iters * 148.0 * 8 * 256;
        printf("mode %d (%s): %.3f ms, %.2f Tflop/s (mul+add counted separately)\n", mode, mode ? "FFMA2 packed" : "FMUL+FADD scalar", ms, flops / ms / 1e9);
    }
    // exactness on 4M random-ish triples incl. signed zeros, denormals, infinities
    const int n = 1 << 22;
    float *a, *b, *c; uint32_t* mism;
    cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&c, n * 4); cudaMallocManaged(&mism, 4);
    uint32_t st = 12345u; *mism = 0;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return st; };
    for (int i = 0; i < n; ++i) {
        uint32_t u[3] = {rnd(), rnd(), rnd()};
        for (int j = 0; j < 3; ++j) { if ((u[j] & 0xFF) == 0) u[j] &= 0x80000000u; if ((u[j] & 0xFF) == 1) u[j] &= 0x807FFFFFu; }
        memcpy(&a[i], &u[0], 4); memcpy(&b[i], &u[1], 4); memcpy(&c[i], &u[2], 4);
        if (a[i] != a[i]) a[i] = 1.5f; if (b[i] != b[i]) b[i] = -2.5f; if (c[i] != c[i]) c[i] = 0.0f;
    }
    exact<<<(n + 255) / 256, 256>>>(a, b, c, mism, n);
    cudaDeviceSynchronize();
    printf("packed vs scalar mismatches (NaN payloads excluded from inputs): %u of %d\n", *mism, n);
    return 0;
}
