// Batch-affine pre-reduction of the sorted bucket entries (stage 4a of the Pippenger pipeline,
// see msm_kernels.cuh).  An affine + affine addition costs 1 inversion + 2M + 1S; with the
// inversions of a whole CTA batch shared through Montgomery's trick that is ~6 field
// multiplications per addition against 10 for the XYZZ mixed addition of k_accumulate.
//
// Round r halves every bucket's slice:   A_r (cnt_r[b] = ceil(n_b / 2^r) points per bucket, dense,
// offsets off_r) -> A_{r+1};  output j of bucket b is  A_r[off_r[b] + 2j] + A_r[off_r[b] + 2j + 1]
// (or a copy of the unpaired last point).  A_0 is virtual: the bases gathered through the sorted
// (index, sign) entries.  Work is split by OUTPUT position, K consecutive outputs per thread, so
// the run time does not depend on the scalar distribution (one hot bucket of 2^20 entries is
// 2^19 independent additions spread over the whole grid).  After R rounds the remaining
// ~n_b / 2^R points per bucket are folded by k_accumulate<DIRECT>.
//
// Per CTA batch (128 threads x K outputs):
//   phase 1  every thread walks its outputs forward, multiplies the denominators (x2 - x1, or 2y
//            for a doubling) into a running product and parks the exclusive prefix in global
//            scratch (coalesced [slot][thread] layout);
//   phase 2  warp shuffles give every lane the product of the lanes before / after it, the four
//            warp totals meet in shared memory, ONE field inversion per CTA;
//   phase 3  every thread walks its outputs backward, peels the individual inverses off the
//            running inverse and finishes the additions.
// Identity operands, P + P and P + (-P) are classified by one shared routine in both phases.
//
// Measured on B200 (2^20 terms, 16 precomputed rows, 4 rounds): 0.35 ns per addition = 60 % of the
// multiplier peak; the operand loads sit on the critical path (long-scoreboard stalls) and are
// covered by occupancy (4 CTAs / SM at 128 registers).  A software-pipelined variant (flag-bit
// walker, cp.async double buffers in shared memory, 4-deep x prefetch) was built and measured:
// 10 % faster on the dense rounds but 30 % slower on the gather round (its deeper queues thrash
// the DRAM random-access stream: 1.5 instead of 2.2 TB/s of 128-byte line fetches), slower overall,
// so it is not kept.
//
// Replaces: the mixed Jacobian additions of ark-ec 0.4 VariableBaseMSM::msm_bigint's bucket loop
// (call sites legogroth16/src/prover.rs:286,299,344,363,592).
#pragma once
#include "ec.cuh"
#include "fp_inv.cuh"

namespace dg {

#define DG_BA_THREADS 128
#define DG_BA_WARPS (DG_BA_THREADS / 32)
#ifndef DG_BA_G2_CTAS
#define DG_BA_G2_CTAS 2          // resident CTAs per SM of the G2 instance (254 registers); 3 (168 registers) spills: measured slower
#endif

// out-of-line multiplier for the once-per-batch phase 2 (keeps the kernel inside the instruction cache)
static __device__ __noinline__ Fp ba_mul(Fp a, Fp b) { return fp_mul(a, b); }
static __device__ __noinline__ Fp2 ba_mul(Fp2 a, Fp2 b) { return fmul(a, b); }
static __device__ __noinline__ Fp ba_inv(Fp a) { return fp_inv_pornin(a); }
static __device__ __noinline__ Fp2 ba_inv(Fp2 a) {
    Fp n = fp_add(fp_mul_ni(a.c0, a.c0), fp_mul_ni(a.c1, a.c1));
    Fp ni = fp_inv_pornin(n);
    return {fp_mul_ni(a.c0, ni), fp_neg(fp_mul_ni(a.c1, ni))};
}

__device__ __forceinline__ Fp ba_shfl_up(const Fp &a, int d) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_up_sync(0xffffffffu, a.l[i], d);
    return r;
}
__device__ __forceinline__ Fp ba_shfl_down(const Fp &a, int d) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_down_sync(0xffffffffu, a.l[i], d);
    return r;
}
__device__ __forceinline__ Fp2 ba_shfl_up(const Fp2 &a, int d) { return {ba_shfl_up(a.c0, d), ba_shfl_up(a.c1, d)}; }
__device__ __forceinline__ Fp2 ba_shfl_down(const Fp2 &a, int d) { return {ba_shfl_down(a.c0, d), ba_shfl_down(a.c1, d)}; }

// Operand pairs that are not a plain chord addition (identity operand, equal x).  Returns true
// when the result is  lambda = num / den  (num, den set; den != 0), false when `direct` already
// is the result.  Called with identical operands from phase 1 and phase 3.
template <class F>
static __device__ __noinline__ bool ba_classify_rare(const Affine<F> &p, const Affine<F> &q, F &num, F &den, Affine<F> &direct) {
    if (aff_is_inf(p)) { direct = q; return false; }
    if (aff_is_inf(q)) { direct = p; return false; }
    if (feq(p.x, q.x)) {
        if (feq(p.y, q.y) && !fis_zero(p.y)) {              // doubling: lambda = 3 x^2 / 2 y
            F xx = ba_mul(p.x, p.x);
            num = fadd(fdbl(xx), xx);
            den = fdbl(p.y);
            return true;
        }
        direct = {fzero<F>(), fzero<F>()};                   // P + (-P)
        return false;
    }
    num = fsub(q.y, p.y);
    den = fsub(q.x, p.x);
    return true;
}

template <class F, bool GATHER>
__device__ __forceinline__ Affine<F> ba_load_point(const Affine<F> *__restrict__ in, const uint32_t *__restrict__ entries, uint32_t i) {
    if (GATHER) {
        uint32_t ent = __ldg(&entries[i]);
        Affine<F> p = aff_load<F>(&in[ent & 0x7fffffffu]);
        p.y = fcneg(p.y, (ent >> 31) != 0);
        return p;
    }
    return aff_load<F>(&in[i]);
}
template <class F, bool GATHER>
__device__ __forceinline__ F ba_load_x(const Affine<F> *__restrict__ in, const uint32_t *__restrict__ entries, uint32_t i) {
    if (GATHER) return fload<F>(&in[__ldg(&entries[i]) & 0x7fffffffu]);
    return fload<F>(&in[i]);
}

// prefix scratch: slot k of thread g lives at pre[(k * NV + j) * nthreads + g], j < NV = sizeof(F) / 16
template <class F> __device__ __forceinline__ void ba_pre_store(uint4 *pre, uint32_t slot, uint32_t nthreads, uint32_t g, const F &v) {
    constexpr int NV = sizeof(F) / 16;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(&v);
#pragma unroll
    for (int j = 0; j < NV; j++)
        pre[((size_t)slot * NV + j) * nthreads + g] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}
template <class F> __device__ __forceinline__ F ba_pre_load(const uint4 *pre, uint32_t slot, uint32_t nthreads, uint32_t g) {
    constexpr int NV = sizeof(F) / 16;
    F v;
    uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
    for (int j = 0; j < NV; j++) {
        uint4 t = pre[((size_t)slot * NV + j) * nthreads + g];
        w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
    }
    return v;
}

// 4 CTAs / SM at 128 registers (G1).  Measured: capping the allocation at 96 registers for 5 CTAs / SM spills ~200 bytes
// per thread and is 4 % slower end to end (7.34 vs 7.04 ms at 2^20 terms).
template <class F, bool GATHER>
__global__ void __launch_bounds__(DG_BA_THREADS, (sizeof(F) > 48 ? DG_BA_G2_CTAS : 4))
    k_affine_round(const Affine<F> *__restrict__ in, const uint32_t *__restrict__ entries, const uint32_t *__restrict__ off_in,
                   const uint32_t *__restrict__ off_out, uint32_t nb, uint32_t K, Affine<F> *__restrict__ out,
                   uint4 *__restrict__ pre) {
    __shared__ F s_wtot[DG_BA_WARPS], s_winv[DG_BA_WARPS];
    const uint32_t nthreads = gridDim.x * blockDim.x;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t Mout = off_out[nb];
    const uint64_t o64 = (uint64_t)gtid * K;
    const uint32_t o0 = o64 < Mout ? (uint32_t)o64 : Mout;
    const uint32_t o1 = (Mout - o0 > K) ? o0 + K : Mout;                 // this thread's outputs [o0, o1)
    uint32_t b = 0;
    if (o0 < o1) {                                                       // off_out[b] <= o0 < off_out[b + 1]
        uint32_t lo = 0, hi = nb;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (off_out[mid] <= o0) lo = mid; else hi = mid;
        }
        b = lo;
    }

    // ---- phase 1: running product of the denominators ------------------------------------------
    // The bucket walk and the entry words of output o + 1 are fetched while output o is processed,
    // so the operand loads of an iteration issue at its top instead of behind two dependent loads.
    constexpr uint32_t MASK = 0x7fffffffu;
    F run = fone<F>();
    if (o0 < o1) {
        uint32_t bo = off_out[b], bo_next = off_out[b + 1], bi = off_in[b], bi_next = off_in[b + 1];
        uint32_t iN = 0, eN0 = 0, eN1 = 0;
        bool pairN = false;
        auto describe = [&](uint32_t o) {
            while (o >= bo_next) {
                b++;
                bo = bo_next; bo_next = off_out[b + 1];
                bi = bi_next; bi_next = off_in[b + 1];
            }
            iN = bi + 2 * (o - bo);
            pairN = iN + 1 < bi_next;
            if (GATHER && pairN) { eN0 = __ldg(&entries[iN]); eN1 = __ldg(&entries[iN + 1]); }
        };
        describe(o0);
        for (uint32_t o = o0; o < o1; o++) {
            const uint32_t i0 = iN, e0 = eN0, e1 = eN1;
            const bool pair = pairN;
            if (o + 1 < o1) describe(o + 1);
            if (!pair) continue;                                         // unpaired last point: copied in phase 3
            F x1 = fload<F>(&in[GATHER ? (e0 & MASK) : i0]), x2 = fload<F>(&in[GATHER ? (e1 & MASK) : i0 + 1]);
            F den;
            bool use = true;
            if (__builtin_expect(feq(x1, x2) || fis_zero(x1) || fis_zero(x2), 0)) {
                Affine<F> p = ba_load_point<F, GATHER>(in, entries, i0), q = ba_load_point<F, GATHER>(in, entries, i0 + 1), direct;
                F num;
                use = ba_classify_rare(p, q, num, den, direct);
            } else {
                den = fsub(x2, x1);
            }
            if (use) {
                ba_pre_store<F>(pre, o - o0, nthreads, gtid, run);
                run = fmul(run, den);
            }
        }
    }

    // ---- phase 2: one inversion per CTA -----------------------------------------------------------
    F incl = run, sincl = run;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        F t = ba_shfl_up(incl, d);
        F m = ba_mul(incl, t);
        incl = fsel(lane >= (uint32_t)d, m, incl);
    }
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        F t = ba_shfl_down(sincl, d);
        F m = ba_mul(sincl, t);
        sincl = fsel(lane + (uint32_t)d < 32u, m, sincl);
    }
    F before = ba_shfl_up(incl, 1), after = ba_shfl_down(sincl, 1);
    if (lane == 0) before = fone<F>();
    if (lane == 31) after = fone<F>();
    if (lane == 31) s_wtot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        F all = s_wtot[0];
#pragma unroll 1
        for (int j = 1; j < DG_BA_WARPS; j++) all = ba_mul(all, s_wtot[j]);
        F oth = ba_inv(all);                                             // same value in every lane: no divergence
#pragma unroll 1
        for (int j = 0; j < DG_BA_WARPS; j++) {                          // lane w: 1 / total of warp w
            F m = ba_mul(oth, s_wtot[j]);
            oth = fsel((uint32_t)j != lane, m, oth);
        }
        if (lane < DG_BA_WARPS) s_winv[lane] = oth;
    }
    __syncthreads();
    F inv_run = ba_mul(ba_mul(s_winv[wid], before), after);            // 1 / (this thread's product)

    // ---- phase 3: peel the inverses off backwards and finish the additions -----------------------
    if (o0 < o1) {
        uint32_t bo = off_out[b], bi = off_in[b], bi_next = off_in[b + 1];
        uint32_t iN = 0, eN0 = 0, eN1 = 0;
        bool pairN = false;
        auto describe = [&](uint32_t o) {
            while (o < bo) {
                b--;
                bo = off_out[b];
                bi_next = bi; bi = off_in[b];
            }
            iN = bi + 2 * (o - bo);
            pairN = iN + 1 < bi_next;
            if (GATHER) {
                eN0 = __ldg(&entries[iN]);
                if (pairN) eN1 = __ldg(&entries[iN + 1]);
            }
        };
        describe(o1 - 1);
        for (uint32_t o = o1; o-- > o0;) {
            const uint32_t i0 = iN, e0 = eN0, e1 = eN1;
            const bool pair = pairN;
            if (o > o0) describe(o - 1);
            Affine<F> p = aff_load<F>(&in[GATHER ? (e0 & MASK) : i0]);
            if (GATHER) p.y = fcneg(p.y, (e0 >> 31) != 0);
            if (!pair) { aff_store(&out[o], p); continue; }
            Affine<F> q = aff_load<F>(&in[GATHER ? (e1 & MASK) : i0 + 1]);
            if (GATHER) q.y = fcneg(q.y, (e1 >> 31) != 0);
            F num, den;
            if (__builtin_expect(feq(p.x, q.x) || fis_zero(p.x) || fis_zero(q.x), 0)) {
                Affine<F> pc = p, qc = q, direct;                        // copies keep p, q out of local memory on the hot path
                F n2, d2;
                if (!ba_classify_rare(pc, qc, n2, d2, direct)) { aff_store(&out[o], direct); continue; }
                num = n2; den = d2;
            } else {
                num = fsub(q.y, p.y);
                den = fsub(q.x, p.x);
            }
            F inv_d = fmul(inv_run, ba_pre_load<F>(pre, o - o0, nthreads, gtid));
            inv_run = fmul(inv_run, den);
            F lam = fmul(num, inv_d);
            Affine<F> r;
            r.x = fsub(fsub(fsqr(lam), p.x), q.x);
            r.y = fsub(fmul(lam, fsub(p.x, r.x)), p.y);
            aff_store(&out[o], r);
        }
    }
}

// cnt_r[b] = ceil(n_b / 2^r) for r = 1 .. R, written to cnt + (r - 1) * stride
static __global__ void __launch_bounds__(256) k_round_counts(const uint32_t *__restrict__ off, uint32_t nb, int R, uint32_t *__restrict__ cnt,
                                                      size_t stride) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t n = off[b + 1] - off[b];
    for (int r = 1; r <= R; r++) cnt[(size_t)(r - 1) * stride + b] = (n + (1u << r) - 1) >> r;
}

}  // namespace dg
