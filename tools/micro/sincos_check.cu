// GPU check: does sincos(double) give bit-identical results to separate sin() and cos() (after rounding to float, as
// libmmd's math:: wrappers do)?  Sweeps every float in [-64, 64] (about 2.2e9 values) and powers-of-two-scaled samples.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t lo, uint32_t n, int negate, unsigned long long* bad_f, unsigned long long* bad_d) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bf = 0, bd = 0;
    for (uint32_t u = lo + i; u < lo + n && u >= lo; u += gridDim.x * blockDim.x) {
        float x = __uint_as_float(u);
        if (negate) x = -x;
        double s, c;
        sincos((double)x, &s, &c);
        const double s2 = sin((double)x), c2 = cos((double)x);
        if (__double_as_longlong(s) != __double_as_longlong(s2) || __double_as_longlong(c) != __double_as_longlong(c2)) ++bd;
        if (__float_as_uint((float)s) != __float_as_uint((float)s2) || __float_as_uint((float)c) != __float_as_uint((float)c2)) ++bf;
    }
    if (bf) atomicAdd(bad_f, bf);
    if (bd) atomicAdd(bad_d, bd);
}
int main() {
    unsigned long long *bf, *bd;
    cudaMallocManaged(&bf, 8); cudaMallocManaged(&bd, 8); *bf = *bd = 0;
    const uint32_t hi = 0x42800000u;  // 64.0f: all non-negative floats below it
    for (int neg = 0; neg < 2; ++neg) {
        for (uint32_t lo = 0; lo < hi; lo += 1u << 28) {
            uint32_t n = (hi - lo < (1u << 28)) ? hi - lo : (1u << 28);
            k<<<148 * 16, 256>>>(lo, n, neg, bf, bd);
        }
    }
    // large arguments (Payne-Hanek path): every 4096th float up to 2^40
    k<<<148 * 16, 256>>>(0x42800000u, 0x53800000u - 0x42800000u, 0, bf, bd);
    cudaDeviceSynchronize();
    printf("sincos vs sin/cos: double-bit mismatches %llu, float-rounded mismatches %llu (%s)\n", *bd, *bf, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
