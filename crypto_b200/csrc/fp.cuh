// BLS12-381 base field Fp on sm_100a: 12 x 32-bit limbs, Montgomery form with R = 2^384
// (bit-identical to ark-ff's Fp<MontBackend<_,6>> when the limbs are read as 6 x u64 LE),
// so coordinates cross the C ABI without conversion.
//
// The multiplier is built for the integer-MAD pipe: every 32x32->64 partial product is a
// mad.lo.cc / madc.hi.cc pair on an even-aligned register pair so ptxas fuses it into one
// IMAD.WIDE.U32(.X) with carry-in/out.  Products whose column index is even go to one
// accumulator, odd columns to a second one ("even/odd split"), which keeps each row a single
// unbroken carry chain; the per-row right shift of the CIOS Montgomery step is absorbed by
// swapping the roles of the two accumulators instead of moving registers.
//
// Replaces: ark-ff Fp arithmetic under every reference call in SURVEY.md section 8a
// (e.g. the mixed additions inside VariableBaseMSM::msm_bigint called at
// legogroth16/src/prover.rs:286).
#pragma once
#include <stdint.h>
#include "bls_consts.cuh"

namespace dg {

struct __align__(16) Fp {
    uint32_t l[12];
};

// p limbs as immediates
#define DG_PL(i) DG_P##i
__device__ __forceinline__ constexpr uint32_t fp_p_limb(int i) {
    constexpr uint32_t P[12] = {DG_P0, DG_P1, DG_P2, DG_P3, DG_P4, DG_P5, DG_P6, DG_P7, DG_P8, DG_P9, DG_P10, DG_P11};
    return P[i];
}

__device__ __forceinline__ Fp fp_zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = 0;
    return r;
}
__device__ __forceinline__ Fp fp_one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = DGC_R_ONE[i];
    return r;
}
__device__ __forceinline__ bool fp_is_zero(const Fp &a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) t |= a.l[i];
    return t == 0;
}
__device__ __forceinline__ bool fp_eq(const Fp &a, const Fp &b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) t |= a.l[i] ^ b.l[i];
    return t == 0;
}

// r = a - p if a >= p else a   (a < 2p, optionally with an extra top carry bit)
__device__ __forceinline__ void fp_final_sub(Fp &a, uint32_t carry = 0) {
    uint32_t t[12], br;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(a.l[0]), "n"(DG_P0));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[1]) : "r"(a.l[1]), "n"(DG_P1));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[2]) : "r"(a.l[2]), "n"(DG_P2));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[3]) : "r"(a.l[3]), "n"(DG_P3));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[4]) : "r"(a.l[4]), "n"(DG_P4));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[5]) : "r"(a.l[5]), "n"(DG_P5));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[6]) : "r"(a.l[6]), "n"(DG_P6));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[7]) : "r"(a.l[7]), "n"(DG_P7));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[8]) : "r"(a.l[8]), "n"(DG_P8));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[9]) : "r"(a.l[9]), "n"(DG_P9));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[10]) : "r"(a.l[10]), "n"(DG_P10));
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[11]) : "r"(a.l[11]), "n"(DG_P11));
    asm volatile("subc.u32 %0, %1, 0;" : "=r"(br) : "r"(carry));
    // br == 0xffffffff when the subtraction borrowed past the carry word -> keep a
    bool keep = (br != 0);
#pragma unroll
    for (int i = 0; i < 12; i++) a.l[i] = keep ? a.l[i] : t[i];
}

__device__ __forceinline__ Fp fp_add(const Fp &a, const Fp &b) {
    Fp r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
    for (int i = 1; i < 12; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
    // p < 2^381 so a + b < 2^382: no carry out of limb 11
    fp_final_sub(r);
    return r;
}

__device__ __forceinline__ Fp fp_sub(const Fp &a, const Fp &b) {
    Fp r;
    uint32_t br;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
    for (int i = 1; i < 12; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(br));
    // add back p masked by the borrow
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r.l[0]) : "r"(br & fp_p_limb(0)));
#pragma unroll
    for (int i = 1; i < 11; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r.l[i]) : "r"(br & fp_p_limb(i)));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(r.l[11]) : "r"(br & fp_p_limb(11)));
    return r;
}

__device__ __forceinline__ Fp fp_neg(const Fp &a) {
    Fp z = fp_zero();
    Fp r = fp_sub(z, a);
    // -0 must stay 0 (fp_sub(0,0) already gives 0)
    return r;
}
__device__ __forceinline__ Fp fp_dbl(const Fp &a) { return fp_add(a, a); }

// conditional negate (branch-free select)
__device__ __forceinline__ Fp fp_cneg(const Fp &a, bool neg) {
    Fp n = fp_neg(a), r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = neg ? n.l[i] : a.l[i];
    return r;
}

// ---- Montgomery multiplication ------------------------------------------------------------
// acc[0..11] += {a[0],a[2],..,a[10]} * b as one carry chain; the carry out is left in CC.
__device__ __forceinline__ void dg_cmad_row(uint32_t *acc, const uint32_t *a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[0]) : "r"(a[0]), "r"(b));
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[1]) : "r"(a[0]), "r"(b));
#pragma unroll
    for (int j = 2; j < 12; j += 2) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j + 1]) : "r"(a[j]), "r"(b));
    }
}
// same with the modulus limbs {p[s], p[s+2], ...} as immediates
template <int S>
__device__ __forceinline__ void dg_cmad_row_p(uint32_t *acc, uint32_t m) {
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[0]) : "r"(m), "r"(fp_p_limb(S)));
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[1]) : "r"(m), "r"(fp_p_limb(S)));
#pragma unroll
    for (int j = 2; j < 12; j += 2) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(m), "r"(fp_p_limb(S + j)));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j + 1]) : "r"(m), "r"(fp_p_limb(S + j)));
    }
}
// acc'[j] = acc[j+2] + {a[0],a[2],..}*b with carry-in from CC (the 2-limb right shift of the
// Montgomery step happens here for free); acc'[10], acc'[11] start from zero.
__device__ __forceinline__ void dg_madc_row_rshift(uint32_t *acc, const uint32_t *a, uint32_t b) {
#pragma unroll
    for (int j = 0; j < 10; j += 2) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(acc[j]) : "r"(a[j]), "r"(b), "r"(acc[j + 2]));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(acc[j + 1]) : "r"(a[j]), "r"(b), "r"(acc[j + 3]));
    }
    asm volatile("madc.lo.cc.u32 %0, %1, %2, 0;" : "=r"(acc[10]) : "r"(a[10]), "r"(b));
    asm volatile("madc.hi.u32 %0, %1, %2, 0;" : "=r"(acc[11]) : "r"(a[10]), "r"(b));
}

// One CIOS row.  T = al + of * 2^32 with `al` holding columns 0..11 and `of` columns 1..12.
// For rows after the first the caller passes the arrays already role-swapped: `al` is the
// previous row's offset array (its columns dropped by one after the division by 2^32) and `of`
// is the previous aligned array, now stale: of[0] == 0, of[1] belongs to column 0, of[2..11]
// to columns 1..10.  Bound: T + a*b_i + m*p < 2^416, and every partial sum is non-negative, so
// neither array overflows (only `al`'s carry out of column 11 is real and lands in of[11]).
template <bool FIRST>
__device__ __forceinline__ void dg_mont_row(uint32_t *al, uint32_t *of, const uint32_t *a, uint32_t bi) {
    if (FIRST) {
#pragma unroll
        for (int j = 0; j < 12; j += 2) {
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(of[j]) : "r"(a[j + 1]), "r"(bi));
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(of[j + 1]) : "r"(a[j + 1]), "r"(bi));
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(al[j]) : "r"(a[j]), "r"(bi));
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(al[j + 1]) : "r"(a[j]), "r"(bi));
        }
    } else {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(al[0]) : "r"(of[1]));   // carry -> column 1
        dg_madc_row_rshift(of, a + 1, bi);                                     // columns 1..12
        dg_cmad_row(al, a, bi);                                                // columns 0..11
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(of[11]));                    // column 12
    }
    uint32_t m = al[0] * DG_FP_INV32;
    dg_cmad_row_p<1>(of, m);
    dg_cmad_row_p<0>(al, m);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(of[11]));
}

__device__ __forceinline__ Fp fp_mul(const Fp &a, const Fp &b) {
    uint32_t ev[12], od[12];
    dg_mont_row<true>(ev, od, a.l, b.l[0]);
#pragma unroll
    for (int i = 1; i < 12; i += 2) {
        dg_mont_row<false>(od, ev, a.l, b.l[i]);
        if (i + 1 < 12) dg_mont_row<false>(ev, od, a.l, b.l[i + 1]);
    }
    // Row 11 ran with al = od, of = ev: the live aligned array is now `ev` (columns 0..11) and
    // `od` is stale with od[0] == 0, od[k] belonging to column k-1.
    Fp r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(ev[0]), "r"(od[1]));
#pragma unroll
    for (int i = 1; i < 11; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(ev[i]), "r"(od[i + 1]));
    asm volatile("addc.u32 %0, %1, 0;" : "=r"(r.l[11]) : "r"(ev[11]));
    fp_final_sub(r);
    return r;
}

__device__ __forceinline__ Fp fp_sqr(const Fp &a) { return fp_mul(a, a); }

}  // namespace dg

// ---- generic field interface (overloaded for Fp and Fp2) so curve code is written once -------
namespace dg {
__device__ __forceinline__ Fp fadd(const Fp &a, const Fp &b) { return fp_add(a, b); }
__device__ __forceinline__ Fp fsub(const Fp &a, const Fp &b) { return fp_sub(a, b); }
__device__ __forceinline__ Fp fmul(const Fp &a, const Fp &b) { return fp_mul(a, b); }
__device__ __forceinline__ Fp fsqr(const Fp &a) { return fp_sqr(a); }
__device__ __forceinline__ Fp fneg(const Fp &a) { return fp_neg(a); }
__device__ __forceinline__ Fp fdbl(const Fp &a) { return fp_dbl(a); }
__device__ __forceinline__ Fp fcneg(const Fp &a, bool n) { return fp_cneg(a, n); }
__device__ __forceinline__ bool fis_zero(const Fp &a) { return fp_is_zero(a); }
__device__ __forceinline__ bool feq(const Fp &a, const Fp &b) { return fp_eq(a, b); }
__device__ __forceinline__ Fp fsel(bool c, const Fp &a, const Fp &b) {   // c ? a : b
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
template <class F> __device__ __forceinline__ F fzero();
template <class F> __device__ __forceinline__ F fone();
template <> __device__ __forceinline__ Fp fzero<Fp>() { return fp_zero(); }
template <> __device__ __forceinline__ Fp fone<Fp>() { return fp_one(); }

// 128-bit vector loads/stores of one field element (48 B, 16-B aligned records)
__device__ __forceinline__ Fp fp_load(const void *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    Fp r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 v = __ldg(q + i);
        r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
    }
    return r;
}
__device__ __forceinline__ Fp fp_load_rw(const void *p) {   // coherent load (data written by earlier kernels/threads)
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    Fp r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 v = q[i];
        r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
    }
    return r;
}
__device__ __forceinline__ void fp_store(void *p, const Fp &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = make_uint4(a.l[4 * i], a.l[4 * i + 1], a.l[4 * i + 2], a.l[4 * i + 3]);
}
}  // namespace dg
