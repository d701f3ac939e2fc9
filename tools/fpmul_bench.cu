// Throughput/latency comparison of the 12x32-bit carry-chain Montgomery multiplier (fp.cuh) and a
// carry-free 14x28-bit variant (full-rate IMAD.WIDE).  nvcc -I crypto_b200/csrc ...
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp.cuh"
using namespace dg;

#define P28(i) DG28_P##i
__device__ __forceinline__ constexpr uint32_t p28(int i) {
    constexpr uint32_t P[14] = {DG28_P0, DG28_P1, DG28_P2, DG28_P3, DG28_P4, DG28_P5, DG28_P6, DG28_P7, DG28_P8, DG28_P9, DG28_P10, DG28_P11, DG28_P12, DG28_P13};
    return P[i];
}
struct F28 { uint32_t l[14]; };
__device__ __forceinline__ F28 mul28(const F28 &a, const F28 &b) {
    uint64_t t[28];
#pragma unroll
    for (int k = 0; k < 28; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) {
#pragma unroll
        for (int j = 0; j < 14; j++) t[i + j] += (uint64_t)a.l[j] * b.l[i];
        uint32_t m = ((uint32_t)t[i] * DG28_PINV) & 0x0fffffffu;
#pragma unroll
        for (int j = 0; j < 14; j++) t[i + j] += (uint64_t)m * p28(j);
        t[i + 1] += t[i] >> 28;
    }
    F28 r;
    uint64_t c = 0;
#pragma unroll
    for (int k = 0; k < 13; k++) { c += t[14 + k]; r.l[k] = (uint32_t)c & 0x0fffffffu; c >>= 28; }
    c += t[27];
    r.l[13] = (uint32_t)c;
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_bench(uint32_t *out, int iters) {
    uint32_t seed = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 0) {
        Fp x, y;
        for (int i = 0; i < 12; i++) { x.l[i] = seed * 2654435761u + i; y.l[i] = seed * 40503u + i * 7; }
        x.l[11] &= 0x0fffffff; y.l[11] &= 0x0fffffff;
        for (int it = 0; it < iters; it++) { x = fp_mul(x, y); y = fp_mul(y, x); }
        uint32_t s = 0; for (int i = 0; i < 12; i++) s ^= x.l[i] ^ y.l[i];
        out[seed] = s;
    } else {
        F28 x, y;
        for (int i = 0; i < 14; i++) { x.l[i] = (seed * 2654435761u + i) & 0x0fffffff; y.l[i] = (seed * 40503u + i * 7) & 0x0fffffff; }
        x.l[13] &= 0xffff; y.l[13] &= 0xffff;
        for (int it = 0; it < iters; it++) { x = mul28(x, y); y = mul28(y, x); }
        uint32_t s = 0; for (int i = 0; i < 14; i++) s ^= x.l[i] ^ y.l[i];
        out[seed] = s;
    }
}

template <int MODE> void run(const char *name, int blocks, int iters, uint32_t *out) {
    k_bench<MODE><<<blocks, 128>>>(out, 4);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_bench<MODE><<<blocks, 128>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mults = (double)blocks * 128 * iters * 2;
    printf("%-28s blocks=%5d  %8.3f ms  %.3e mults/s  (%.0f cycles/mult/thread-serial at 1.965GHz)\n", name, blocks, ms, mults / (ms * 1e-3),
           ms * 1e-3 * 1.965e9 / (iters * 2));
}
int main() {
    uint32_t *out; cudaMalloc(&out, 148 * 32 * 128 * 4);
    for (int bpsm : {1, 2, 3, 4, 8}) {
        run<0>("12x32 carry-chain", 148 * bpsm, 2000, out);
        run<1>("14x28 carry-free", 148 * bpsm, 2000, out);
    }
    run<0>("12x32 single warp/SM", 148, 2000, out);
    return 0;
}
