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

// ---- dedicated squaring ----------------------------------------------------------------------
// a^2 = sum_i a_i^2 2^(64 i) + 2 sum_{i<j} a_i a_j 2^(32 (i + j)): 66 cross products + 12 squares instead of 144 products,
// then the 12 reduction rows of the multiplier on the low half of the 24-limb square and one addition of the high half.
// 222 wide multiply-adds against 288 for fp_mul(a, a).  The cross products keep the even/odd split: column i + j even goes
// to A (A[k] = column k), odd to O (O[k] = column k + 1), so every product is one IMAD.WIDE on an aligned register pair and
// each row is two unbroken carry chains whose carry-out falls into a limb no earlier row has touched.
//
// acc'[j] = acc[j + 2] + {p[1], p[3], ..} * m with carry-in from CC: the reduction row's half that also absorbs the shift
__device__ __forceinline__ void dg_madc_row_rshift_p(uint32_t *acc, uint32_t m) {
#pragma unroll
    for (int j = 0; j < 10; j += 2) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(acc[j]) : "r"(m), "r"(fp_p_limb(1 + j)), "r"(acc[j + 2]));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(acc[j + 1]) : "r"(m), "r"(fp_p_limb(1 + j)), "r"(acc[j + 3]));
    }
    asm volatile("madc.lo.cc.u32 %0, %1, %2, 0;" : "=r"(acc[10]) : "r"(m), "r"(fp_p_limb(11)));
    asm volatile("madc.hi.u32 %0, %1, %2, 0;" : "=r"(acc[11]) : "r"(m), "r"(fp_p_limb(11)));
}
// one Montgomery reduction row without a product row: (al, of) as in dg_mont_row<false>
__device__ __forceinline__ void dg_redc_row(uint32_t *al, uint32_t *of) {
    const uint32_t m = (al[0] + of[1]) * DG_FP_INV32;
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(al[0]) : "r"(of[1]));   // carry -> column 1
    dg_madc_row_rshift_p(of, m);                                           // columns 1..12
    dg_cmad_row_p<0>(al, m);                                               // columns 0..11
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(of[11]));                    // column 12
}

// T / 2^384 mod p for a 24-limb T < p * 2^384 (T is clobbered): the 12 reduction rows of the multiplier on the low half,
// roles swapped as in fp_mul, then one addition of the high half.  (T + M p) / R < T / R + p < 2p: one final subtraction.
__device__ __forceinline__ Fp fp_redc24(uint32_t *T) {
    uint32_t od[12];
#pragma unroll
    for (int k = 0; k < 12; k++) od[k] = 0;
#pragma unroll
    for (int i = 0; i < 12; i += 2) {
        dg_redc_row(T, od);
        dg_redc_row(od, T);
    }
    // live aligned array: T[0..11]; od is stale (od[k] belongs to column k - 1)
    Fp r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(T[0]), "r"(od[1]));
#pragma unroll
    for (int i = 1; i < 11; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(T[i]), "r"(od[i + 1]));
    asm volatile("addc.u32 %0, %1, 0;" : "=r"(r.l[11]) : "r"(T[11]));
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r.l[0]) : "r"(T[12]));
#pragma unroll
    for (int i = 1; i < 11; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r.l[i]) : "r"(T[12 + i]));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(r.l[11]) : "r"(T[23]));
    fp_final_sub(r);
    return r;
}

// six products {a[s], a[s + 2], ..} * b accumulated on acc[start .. start + 11] as one carry chain; the carry out lands in
// acc[start + 12], a limb that holds nothing but earlier carries
__device__ __forceinline__ void dg_wide_chain(uint32_t *acc, int start, const uint32_t *a, int s, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[start]) : "r"(a[s]), "r"(b));
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[start + 1]) : "r"(a[s]), "r"(b));
#pragma unroll
    for (int t = 1; t < 6; t++) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[start + 2 * t]) : "r"(a[s + 2 * t]), "r"(b));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[start + 2 * t + 1]) : "r"(a[s + 2 * t]), "r"(b));
    }
    if (start + 12 < 24) asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[start + 12]));
}
// T[0..23] = a * b as a plain integer (a, b < 2^384, a * b < 2^768): the 144 products of the multiplier without the
// interleaved reduction rows, even/odd split over two 24-limb accumulators (A[k] = column k, O[k] = column k + 1)
__device__ __forceinline__ void fp_mul_wide(uint32_t *T, const uint32_t *a, const uint32_t *b) {
    uint32_t O[24];
#pragma unroll
    for (int j = 0; j < 12; j += 2) {
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(T[j]) : "r"(a[j]), "r"(b[0]));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(T[j + 1]) : "r"(a[j]), "r"(b[0]));
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(O[j]) : "r"(a[j + 1]), "r"(b[0]));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(O[j + 1]) : "r"(a[j + 1]), "r"(b[0]));
    }
#pragma unroll
    for (int k = 12; k < 24; k++) { T[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 1; i < 12; i++) {
        if (i & 1) {
            dg_wide_chain(O, i - 1, a, 0, b[i]);          // a_even * b_i: odd columns i + j -> O[i + j - 1]
            dg_wide_chain(T, i + 1, a, 1, b[i]);          // a_odd * b_i: even columns i + j -> A[i + j]
        } else {
            dg_wide_chain(T, i, a, 0, b[i]);
            dg_wide_chain(O, i, a, 1, b[i]);
        }
    }
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(T[1]) : "r"(O[0]));
#pragma unroll
    for (int k = 2; k < 23; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(T[k]) : "r"(O[k - 1]));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(T[23]) : "r"(O[22]));
}
// 24-limb helpers for lazily reduced sums of products
__device__ __forceinline__ void dg_sub24(uint32_t *a, const uint32_t *b) {                 // a -= b  (mod 2^768)
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(b[0]));
#pragma unroll
    for (int k = 1; k < 23; k++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b[k]));
    asm volatile("subc.u32 %0, %0, %1;" : "+r"(a[23]) : "r"(b[23]));
}
__device__ __forceinline__ void dg_add24_psq(uint32_t *a) {                                 // a += p^2  (mod 2^768)
    constexpr uint32_t Q[24] = {DG_PSQ0, DG_PSQ1, DG_PSQ2, DG_PSQ3, DG_PSQ4, DG_PSQ5, DG_PSQ6, DG_PSQ7, DG_PSQ8, DG_PSQ9, DG_PSQ10, DG_PSQ11,
                                DG_PSQ12, DG_PSQ13, DG_PSQ14, DG_PSQ15, DG_PSQ16, DG_PSQ17, DG_PSQ18, DG_PSQ19, DG_PSQ20, DG_PSQ21, DG_PSQ22, DG_PSQ23};
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(Q[0]));
#pragma unroll
    for (int k = 1; k < 23; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(Q[k]));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(a[23]) : "r"(Q[23]));
}
// a + b as a plain integer (no reduction; a, b < p gives < 2p < 2^382)
__device__ __forceinline__ Fp fp_add_raw(const Fp &a, const Fp &b) {
    Fp r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
    for (int i = 1; i < 11; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(r.l[11]) : "r"(a.l[11]), "r"(b.l[11]));
    return r;
}

__device__ __forceinline__ Fp fp_sqr(const Fp &x) {
#ifdef DG_FP_SQR_VIA_MUL
    return fp_mul(x, x);
#else
    const uint32_t *a = x.l;
    uint32_t A[24], O[24];
    // row 0 initialises O[0..11] and A[2..11]
#pragma unroll
    for (int j = 1; j < 12; j += 2) {
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(O[j - 1]) : "r"(a[0]), "r"(a[j]));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(O[j]) : "r"(a[0]), "r"(a[j]));
    }
#pragma unroll
    for (int j = 2; j < 12; j += 2) {
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(A[j]) : "r"(a[0]), "r"(a[j]));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(A[j + 1]) : "r"(a[0]), "r"(a[j]));
    }
#pragma unroll
    for (int k = 12; k < 24; k++) { A[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 1; i < 11; i++) {
        {   // i + j odd -> O[i + j - 1], O[i + j]
            int last = 0;
#pragma unroll
            for (int j = i + 1; j < 12; j += 2) {
                const int k = i + j - 1;
                if (j == i + 1) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(O[k]) : "r"(a[i]), "r"(a[j]));
                else asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(O[k]) : "r"(a[i]), "r"(a[j]));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(O[k + 1]) : "r"(a[i]), "r"(a[j]));
                last = k + 2;
            }
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(O[last]));
        }
        if (i + 2 < 12) {   // i + j even -> A[i + j], A[i + j + 1]
            int last = 0;
#pragma unroll
            for (int j = i + 2; j < 12; j += 2) {
                const int k = i + j;
                if (j == i + 2) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(A[k]) : "r"(a[i]), "r"(a[j]));
                else asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(A[k]) : "r"(a[i]), "r"(a[j]));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(A[k + 1]) : "r"(a[i]), "r"(a[j]));
                last = k + 2;
            }
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(A[last]));
        }
    }
    // T = 2 * (A + (O << 32)) + sum_i a_i^2 2^(64 i);  columns 0 and 1 of the cross products are 0 and O[0]
    uint32_t T[24];
    T[1] = O[0];
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(T[2]) : "r"(A[2]), "r"(O[1]));
#pragma unroll
    for (int k = 3; k < 23; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(T[k]) : "r"(A[k]), "r"(O[k - 1]));
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(T[23]) : "r"(A[23]), "r"(O[22]));
#pragma unroll
    for (int k = 23; k >= 2; k--) T[k] = __funnelshift_l(T[k - 1], T[k], 1);
    T[1] <<= 1;
    T[0] = 0;
    asm volatile("mad.lo.cc.u32 %0, %1, %1, %0;" : "+r"(T[0]) : "r"(a[0]));
    asm volatile("madc.hi.cc.u32 %0, %1, %1, %0;" : "+r"(T[1]) : "r"(a[0]));
#pragma unroll
    for (int i = 1; i < 12; i++) {
        asm volatile("madc.lo.cc.u32 %0, %1, %1, %0;" : "+r"(T[2 * i]) : "r"(a[i]));
        if (i < 11) asm volatile("madc.hi.cc.u32 %0, %1, %1, %0;" : "+r"(T[2 * i + 1]) : "r"(a[i]));
        else asm volatile("madc.hi.u32 %0, %1, %1, %0;" : "+r"(T[2 * i + 1]) : "r"(a[i]));
    }
    return fp_redc24(T);
#endif
}

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
