// Fp2 = Fp[u]/(u^2 + 1) for BLS12-381 G2 and the pairing tower (ark-bls12-381 Fq2, not vendored
// in /root/reference).  Karatsuba multiplication (3 Fp mults) and complex squaring (2 Fp mults).
#pragma once
#include "fp.cuh"

namespace dg {

struct __align__(16) Fp2 {
    Fp c0, c1;
};

__device__ __forceinline__ Fp2 fadd(const Fp2 &a, const Fp2 &b) { return {fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
__device__ __forceinline__ Fp2 fsub(const Fp2 &a, const Fp2 &b) { return {fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
__device__ __forceinline__ Fp2 fneg(const Fp2 &a) { return {fp_neg(a.c0), fp_neg(a.c1)}; }
__device__ __forceinline__ Fp2 fdbl(const Fp2 &a) { return {fp_dbl(a.c0), fp_dbl(a.c1)}; }
__device__ __forceinline__ Fp2 fcneg(const Fp2 &a, bool n) { return {fp_cneg(a.c0, n), fp_cneg(a.c1, n)}; }
__device__ __forceinline__ bool fis_zero(const Fp2 &a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
__device__ __forceinline__ bool feq(const Fp2 &a, const Fp2 &b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
__device__ __forceinline__ Fp2 fsel(bool c, const Fp2 &a, const Fp2 &b) { return {fsel(c, a.c0, b.c0), fsel(c, a.c1, b.c1)}; }
template <> __device__ __forceinline__ Fp2 fzero<Fp2>() { return {fp_zero(), fp_zero()}; }
template <> __device__ __forceinline__ Fp2 fone<Fp2>() { return {fp_one(), fp_zero()}; }

// The Fp multiplier is ~400 SASS instructions; Fp2 and everything above it call it out of line
// so G2 / pairing kernels stay inside the instruction cache.
// The operands are passed BY VALUE (24 registers under the device ABI).  With `const Fp &` parameters the compiler must
// assume they alias the returned object, re-reads the limbs from local memory between the asm statements of the
// multiplier, and ptxas then no longer fuses the mad.lo.cc / madc.hi.cc pairs of the a * b rows into IMAD.WIDE:
// 807 instead of ~500 instructions per call (ncu source counters of k_affine_round<Fp2>, profiles/ncu_affine_round_g2_r02.json).
// Every out-of-line function below follows the same rule: operands by value or copied into locals first.
static __device__ __noinline__ Fp fp_mul_ni(Fp a, Fp b) { return fp_mul(a, b); }

// Karatsuba with lazy reduction: the three 768-bit products a0 b0, a1 b1, (a0 + a1)(b0 + b1) are combined as plain
// integers and only the two results are reduced -- 3 x 144 + 2 x 156 = 744 wide multiply-adds instead of 3 x 300 = 900.
//   c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1 = a0 b1 + a1 b0  in [0, 2 p^2)
//   c0 = a0 b0 - a1 b1 + p^2                                  in (0, 2 p^2)
// both below p * 2^384 (2p < 2^384), so one Montgomery reduction each gives a value below 2p and fp_redc24's single final
// subtraction makes it canonical.  Operands by value: see fp_mul_ni.
#ifdef DG_FP2_MUL_LAZY
static __device__ __noinline__ Fp2 fp2_mul_ni(Fp2 a, Fp2 b) {
    uint32_t t0[24], t1[24], t2[24];
    fp_mul_wide(t0, a.c0.l, b.c0.l);
    fp_mul_wide(t1, a.c1.l, b.c1.l);
    {
        Fp sa = fp_add_raw(a.c0, a.c1), sb = fp_add_raw(b.c0, b.c1);
        fp_mul_wide(t2, sa.l, sb.l);
    }
    dg_sub24(t2, t0);
    dg_sub24(t2, t1);
    dg_sub24(t0, t1);
    dg_add24_psq(t0);
    Fp2 r;
    r.c0 = fp_redc24(t0);
    r.c1 = fp_redc24(t2);
    return r;
}
#endif
// Measured on B200 (tools/ab_bench.py, G2 MSM at 2^18 terms): the lazy form is SLOWER than three out-of-line
// multiplications -- 7.60 vs 7.41 ms on raw bases, 5.39 vs 5.30 ms through a table.  It saves 17 % of the multiply-adds
// but needs 132 registers of its own (the batch-affine kernel around it already sits at the 254-register cap and spills
// ~200 bytes more) and adds ~170 dependent 24-limb carry-chain additions.  Kept for the record behind DG_FP2_MUL_LAZY.
#ifdef DG_FP2_MUL_LAZY
__device__ __forceinline__ Fp2 fmul(const Fp2 &a, const Fp2 &b) { return fp2_mul_ni(a, b); }
#else
__device__ __forceinline__ Fp2 fmul(const Fp2 &a, const Fp2 &b) {
    Fp t0 = fp_mul_ni(a.c0, b.c0);
    Fp t1 = fp_mul_ni(a.c1, b.c1);
    Fp m = fp_mul_ni(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
    Fp2 r;
    r.c0 = fp_sub(t0, t1);
    r.c1 = fp_sub(fp_sub(m, t0), t1);
    return r;
}
#endif
__device__ __forceinline__ Fp2 fsqr(const Fp2 &a) {
    Fp s = fp_add(a.c0, a.c1), d = fp_sub(a.c0, a.c1);
    Fp m = fp_mul_ni(a.c0, a.c1);
    Fp2 r;
    r.c0 = fp_mul_ni(s, d);
    r.c1 = fp_dbl(m);
    return r;
}
__device__ __forceinline__ Fp2 fp2_mul_fp(const Fp2 &a, const Fp &k) { return {fp_mul_ni(a.c0, k), fp_mul_ni(a.c1, k)}; }
__device__ __forceinline__ Fp2 fp2_conj(const Fp2 &a) { return {a.c0, fp_neg(a.c1)}; }
__device__ __forceinline__ Fp2 fp2_mul_xi(const Fp2 &a) { return {fp_sub(a.c0, a.c1), fp_add(a.c0, a.c1)}; }   // * (1 + u)

__device__ __forceinline__ Fp2 fp2_load(const void *p) {
    return {fp_load(p), fp_load(reinterpret_cast<const char *>(p) + 48)};
}
__device__ __forceinline__ Fp2 fp2_load_rw(const void *p) {
    return {fp_load_rw(p), fp_load_rw(reinterpret_cast<const char *>(p) + 48)};
}
__device__ __forceinline__ void fp2_store(void *p, const Fp2 &a) {
    fp_store(p, a.c0);
    fp_store(reinterpret_cast<char *>(p) + 48, a.c1);
}

// element-size-generic load/store used by the templated curve kernels
template <class F> __device__ __forceinline__ F fload(const void *p);
template <class F> __device__ __forceinline__ F fload_rw(const void *p);
template <> __device__ __forceinline__ Fp fload<Fp>(const void *p) { return fp_load(p); }
template <> __device__ __forceinline__ Fp2 fload<Fp2>(const void *p) { return fp2_load(p); }
template <> __device__ __forceinline__ Fp fload_rw<Fp>(const void *p) { return fp_load_rw(p); }
template <> __device__ __forceinline__ Fp2 fload_rw<Fp2>(const void *p) { return fp2_load_rw(p); }
__device__ __forceinline__ void fstore(void *p, const Fp &a) { fp_store(p, a); }
__device__ __forceinline__ void fstore(void *p, const Fp2 &a) { fp2_store(p, a); }

}  // namespace dg
