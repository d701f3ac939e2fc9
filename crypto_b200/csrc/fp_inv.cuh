// Fp inversion by the binary extended Euclidean algorithm (right-shift variant): ~2*381 rounds of
// 12-limb shifts / additions instead of the 476 Montgomery multiplications of a^(p-2).  Used where
// an inversion sits on a serial path: the final exponentiation of the pairing (one per product),
// normalize_batch (one per 16-point chunk), fixed-base table construction.
// Montgomery aware: in = a*R, out = a^-1 * R.
#pragma once
#include "fp.cuh"

namespace dg {

__device__ __forceinline__ bool fp_raw_geq(const Fp &a, const Fp &b) {     // a >= b as integers
    uint32_t br;
    uint32_t t;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(t) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
    for (int i = 1; i < 12; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t) : "r"(a.l[i]), "r"(b.l[i]));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(br));
    return br == 0;
}
__device__ __forceinline__ Fp fp_raw_sub(const Fp &a, const Fp &b) {       // a - b, caller guarantees a >= b
    Fp r;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
    for (int i = 1; i < 11; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
    asm volatile("subc.u32 %0, %1, %2;" : "=r"(r.l[11]) : "r"(a.l[11]), "r"(b.l[11]));
    return r;
}
__device__ __forceinline__ Fp fp_raw_shr1(const Fp &a) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 11; i++) r.l[i] = (a.l[i] >> 1) | (a.l[i + 1] << 31);
    r.l[11] = a.l[11] >> 1;
    return r;
}
// x / 2 mod p for x < p: add p when odd (x + p < 2^382 fits 12 limbs), then shift
__device__ __forceinline__ Fp fp_halve(const Fp &a) {
    uint32_t odd = 0u - (a.l[0] & 1u);
    Fp t;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(t.l[0]) : "r"(a.l[0]), "r"(odd & fp_p_limb(0)));
#pragma unroll
    for (int i = 1; i < 11; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(t.l[i]) : "r"(a.l[i]), "r"(odd & fp_p_limb(i)));
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(t.l[11]) : "r"(a.l[11]), "r"(odd & fp_p_limb(11)));
    return fp_raw_shr1(t);
}
__device__ __forceinline__ bool fp_raw_is_one(const Fp &a) {
    uint32_t t = a.l[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 12; i++) t |= a.l[i];
    return t == 0;
}

static __device__ __noinline__ Fp fp_inv_binary(const Fp &a_mont) {
    if (fp_is_zero(a_mont)) return fp_zero();           // inv(0) = 0 like the Fermat version
    Fp u = a_mont, v, x1 = fp_zero(), x2 = fp_zero();
#pragma unroll
    for (int i = 0; i < 12; i++) v.l[i] = fp_p_limb(i);
    x1.l[0] = 1;
    while (!fp_raw_is_one(u) && !fp_raw_is_one(v)) {
        while (!(u.l[0] & 1)) { u = fp_raw_shr1(u); x1 = fp_halve(x1); }
        while (!(v.l[0] & 1)) { v = fp_raw_shr1(v); x2 = fp_halve(x2); }
        if (fp_raw_geq(u, v)) { u = fp_raw_sub(u, v); x1 = fp_sub(x1, x2); }
        else { v = fp_raw_sub(v, u); x2 = fp_sub(x2, x1); }
    }
    Fp x = fp_raw_is_one(u) ? x1 : x2;                   // x = (a R)^-1 = a^-1 R^-1 as an integer mod p
    Fp r2;
#pragma unroll
    for (int i = 0; i < 12; i++) r2.l[i] = DGC_R2[i];
    x = fp_mul(x, r2);                                   // a^-1 R^-1 * R^2 / R = a^-1
    return fp_mul(x, r2);                                // a^-1 * R^2 / R = a^-1 R
}

}  // namespace dg
