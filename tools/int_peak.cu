// Integer-pipe peak micro-benchmark for the MSM roofline.
// Measures, on the whole chip, the sustained rate of the instructions the Fp multiplier is
// made of: IMAD (mad.lo), IMAD.HI, IMAD.WIDE.U32, the carry-chained IMAD.WIDE.U32.X
// (mad.lo.cc/madc.hi.cc pairs), IADD3.X chains, DFMA, and IMAD+IADD3 co-issue.
// Prints one JSON object; bench.py reads profiles/int_peak_*.json for its integer roofline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak tools/int_peak.cu && ./int_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void __launch_bounds__(256) k_peak(uint32_t *out, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t r[2 * ILP];
    double d[ILP];
    uint64_t w[ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) r[i] = a + i;
#pragma unroll
    for (int i = 0; i < ILP; i++) { d[i] = (double)(a + i); w[i] = ((uint64_t)(a + i) << 32) | (b + i); }
    double da = (double)a * 1e-3, db = (double)b * 1e-3;
    for (int it = 0; it < ITERS; it++) {
        // The multiplicand of every product is a neighbouring accumulator: with loop-invariant operands ptxas computes
        // a * b once and turns the loop into additions (round 1's imad_wide figure was such an IADD3 rate).
        if (MODE == 0) {          // IMAD lo
#pragma unroll
            for (int i = 0; i < 2 * ILP; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(r[(i + 1) % (2 * ILP)]), "r"(b));
        } else if (MODE == 1) {   // IMAD.HI
#pragma unroll
            for (int i = 0; i < 2 * ILP; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(r[(i + 1) % (2 * ILP)]), "r"(b));
        } else if (MODE == 2) {   // IMAD.WIDE.U32 (64-bit accumulate, no carry)
#pragma unroll
            for (int i = 0; i < ILP; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 1) % ILP]), "r"(b));
        } else if (MODE == 3) {   // carry-chained wide MAD: one chain of ILP pairs
            uint32_t m[ILP];
#pragma unroll
            for (int i = 0; i < ILP; i++) m[i] = r[(2 * i + 2) % (2 * ILP)] ^ (uint32_t)it;
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r[0]) : "r"(m[0]), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r[1]) : "r"(m[0]), "r"(b));
#pragma unroll
            for (int i = 1; i < ILP; i++) {
                asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r[2 * i]) : "r"(m[i]), "r"(b));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r[2 * i + 1]) : "r"(m[i]), "r"(b));
            }
        } else if (MODE == 4) {   // IADD3.X chain
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r[0]) : "r"(a));
#pragma unroll
            for (int i = 1; i < 2 * ILP; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
        } else if (MODE == 5) {   // DFMA
#pragma unroll
            for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[(i + 1) % ILP]), "d"(db));
#pragma unroll
            for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[(i + 1) % ILP]), "d"(da));
        } else if (MODE == 6) {   // IMAD + IADD3 co-issue (one each)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(r[ILP + i]) : "r"(b));
            }
        } else if (MODE == 8) {   // IMAD.WIDE and DFMA interleaved one to one: do the two pipes issue side by side?
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 1) % ILP]), "r"(b));
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[(i + 1) % ILP]), "d"(db));
            }
        } else if (MODE == 9) {   // carry-chained wide MAD pairs (the multiplier's rows) interleaved with DFMA
            uint32_t m[ILP];                       // multiplicands of this row, read before the row overwrites them
#pragma unroll
            for (int i = 0; i < ILP; i++) m[i] = r[(2 * i + 2) % (2 * ILP)] ^ (uint32_t)it;
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r[0]) : "r"(m[0]), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r[1]) : "r"(m[0]), "r"(b));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[0]) : "d"(d[1]), "d"(db));
#pragma unroll
            for (int i = 1; i < ILP; i++) {
                asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r[2 * i]) : "r"(m[i]), "r"(b));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r[2 * i + 1]) : "r"(m[i]), "r"(b));
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[(i + 1) % ILP]), "d"(db));
            }
        } else if (MODE == 7) {   // split form ptxas prefers: IMAD + IMAD.HI + 2 x IADD3.X per product
            asm volatile("add.cc.u32 %0, %0, 0;" : "+r"(r[0]));
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                uint32_t lo, hi;
                asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(a + i), "r"(b));
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(a + i), "r"(b));
                asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r[2 * i]) : "r"(lo));
                asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r[2 * i + 1]) : "r"(hi));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) s ^= r[i];
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= (uint32_t)d[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (s == 0x12345678) out[0] = s;
}

template <int MODE>
double run(const char *name, double ops_per_iter_per_thread, uint32_t *out, int nsm) {
    int blocks = nsm * 8, threads = 256;
    k_peak<MODE><<<blocks, threads>>>(out, 7);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_peak<MODE><<<blocks, threads>>>(out, 7 + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * ITERS * ops_per_iter_per_thread;
    double rate = ops / (best * 1e-3);
    printf("  \"%s\": {\"ops_per_s\": %.4e, \"per_sm_per_clk_at_1965MHz\": %.2f, \"ms\": %.3f},\n", name, rate,
           rate / nsm / 1.965e9, best);
    return rate;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    uint32_t *out;
    cudaMalloc(&out, 64);
    int nsm = prop.multiProcessorCount;
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d,\n", prop.name, nsm);
    run<0>("imad_lo", 2 * ILP, out, nsm);
    run<1>("imad_hi", 2 * ILP, out, nsm);
    run<2>("imad_wide", ILP, out, nsm);
    run<3>("imad_wide_x_chain", ILP, out, nsm);
    run<4>("iadd3_x_chain", 2 * ILP, out, nsm);
    run<5>("dfma", 2 * ILP, out, nsm);
    run<6>("imad_plus_iadd_pairs", ILP, out, nsm);
    run<7>("split_mul_lo_hi_addc_products", ILP, out, nsm);
    run<8>("imad_wide_and_dfma_interleaved", 2 * ILP, out, nsm);
    run<9>("imad_wide_x_chain_and_dfma_interleaved", 2 * ILP, out, nsm);
    printf("  \"note\": \"ops = instructions of the named kind (pairs/products for split_mul / x_chain; wide MACs + DFMAs together for the interleaved modes), whole chip\"\n}\n");
    return 0;
}
