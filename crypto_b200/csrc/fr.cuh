// BLS12-381 scalar field Fr on sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256
// (bit-identical to ark-ff's Fr when read as 4 x u64 LE).  Same even/odd-column CIOS multiplier
// as fp.cuh, specialised to 8 limbs.  Used by the NTT.
#pragma once
#include <stdint.h>
#include "bls_consts.cuh"

namespace dg {

struct __align__(16) Fr {
    uint32_t l[8];
};

__device__ __forceinline__ constexpr uint32_t fr_mod_limb(int i) {
    constexpr uint32_t M[8] = {DG_R0, DG_R1, DG_R2, DG_R3, DG_R4, DG_R5, DG_R6, DG_R7};
    return M[i];
}
__device__ __forceinline__ Fr fr_zero() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
}
__device__ __forceinline__ Fr fr_const(const uint32_t *c) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c[i];
    return r;
}
__device__ __forceinline__ Fr fr_one() { return fr_const(DGC_FR_ONE); }

// a - r if a >= r (a < 2r, optional carry word)
__device__ __forceinline__ void fr_final_sub(Fr &a, uint32_t carry = 0) {
    uint32_t t[8], br;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(a.l[0]), "r"(fr_mod_limb(0)));
#pragma unroll
    for (int i = 1; i < 8; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[i]) : "r"(a.l[i]), "r"(fr_mod_limb(i)));
    asm volatile("subc.u32 %0, %1, 0;" : "=r"(br) : "r"(carry));
    bool keep = (br != 0);
#pragma unroll
    for (int i = 0; i < 8; i++) a.l[i] = keep ? a.l[i] : t[i];
}
__device__ __forceinline__ Fr fr_add(const Fr &a, const Fr &b) {
    Fr r;
    uint32_t c;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
    for (int i = 1; i < 8; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));      // r < 2^255 so the sum fits 256 bits: c == 0, kept for safety
    fr_final_sub(r, c);
    return r;
}
__device__ __forceinline__ Fr fr_sub(const Fr &a, const Fr &b) {
    Fr r;
    uint32_t br;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
    for (int i = 1; i < 8; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(br));
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r.l[0]) : "r"(br & fr_mod_limb(0)));
#pragma unroll
    for (int i = 1; i < 7; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r.l[i]) : "r"(br & fr_mod_limb(i)));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(r.l[7]) : "r"(br & fr_mod_limb(7)));
    return r;
}

// ---- Montgomery multiplication, 8 limbs (see fp.cuh for the scheme) --------------------------
__device__ __forceinline__ void fr_cmad_row(uint32_t *acc, const uint32_t *a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[0]) : "r"(a[0]), "r"(b));
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[1]) : "r"(a[0]), "r"(b));
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j + 1]) : "r"(a[j]), "r"(b));
    }
}
template <int S> __device__ __forceinline__ void fr_cmad_row_m(uint32_t *acc, uint32_t m) {
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[0]) : "r"(m), "r"(fr_mod_limb(S)));
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[1]) : "r"(m), "r"(fr_mod_limb(S)));
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(m), "r"(fr_mod_limb(S + j)));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j + 1]) : "r"(m), "r"(fr_mod_limb(S + j)));
    }
}
__device__ __forceinline__ void fr_madc_row_rshift(uint32_t *acc, const uint32_t *a, uint32_t b) {
#pragma unroll
    for (int j = 0; j < 6; j += 2) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(acc[j]) : "r"(a[j]), "r"(b), "r"(acc[j + 2]));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(acc[j + 1]) : "r"(a[j]), "r"(b), "r"(acc[j + 3]));
    }
    asm volatile("madc.lo.cc.u32 %0, %1, %2, 0;" : "=r"(acc[6]) : "r"(a[6]), "r"(b));
    asm volatile("madc.hi.u32 %0, %1, %2, 0;" : "=r"(acc[7]) : "r"(a[6]), "r"(b));
}
template <bool FIRST> __device__ __forceinline__ void fr_mont_row(uint32_t *al, uint32_t *of, const uint32_t *a, uint32_t bi) {
    if (FIRST) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(of[j]) : "r"(a[j + 1]), "r"(bi));
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(of[j + 1]) : "r"(a[j + 1]), "r"(bi));
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(al[j]) : "r"(a[j]), "r"(bi));
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(al[j + 1]) : "r"(a[j]), "r"(bi));
        }
    } else {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(al[0]) : "r"(of[1]));
        fr_madc_row_rshift(of, a + 1, bi);
        fr_cmad_row(al, a, bi);
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(of[7]));
    }
    uint32_t m = al[0] * DG_FR_INV32;
    fr_cmad_row_m<1>(of, m);
    fr_cmad_row_m<0>(al, m);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(of[7]));
}
// r < 2^255 leaves only one spare bit in 256: the running sum T + a*b_i + m*r stays below
// 2^32 * 2r + ... < 2^289, i.e. inside the 9 columns (288 bits) only because a, b < r < 2^255:
// a*b_i + m*r + T < 2^287 + 2^287 + 2^256 < 2^288.
__device__ __forceinline__ Fr fr_mul(const Fr &a, const Fr &b) {
    uint32_t ev[8], od[8];
    fr_mont_row<true>(ev, od, a.l, b.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        fr_mont_row<false>(od, ev, a.l, b.l[i]);
        if (i + 1 < 8) fr_mont_row<false>(ev, od, a.l, b.l[i + 1]);
    }
    Fr r;
    uint32_t c;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(ev[0]), "r"(od[1]));
#pragma unroll
    for (int i = 1; i < 7; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(ev[i]), "r"(od[i + 1]));
    asm volatile("addc.cc.u32 %0, %1, 0;" : "=r"(r.l[7]) : "r"(ev[7]));
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
    fr_final_sub(r, c);
    return r;
}

__device__ __forceinline__ Fr fr_load(const void *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void fr_store(void *p, const Fr &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
    q[1] = make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]);
}

}  // namespace dg
