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

static __device__ __noinline__ Fp fp_inv_binary(Fp a_mont) {
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

// ---- Pornin's optimised binary GCD (eprint 2020/972, algorithm 2, k = 32) ---------------------------------------------
// 25 outer rounds; each runs 31 divsteps on 64-bit approximations of (a, b) (the low 31 bits exactly, the top 33 bits
// of the longer one) collecting a 2x2 transition matrix with entries |f|, |g| <= 2^31, then applies the matrix once to
// the 381-bit values (a, b) and -- with a 31-bit Montgomery-style division -- to the Bezout coefficients (u, v) kept in
// [0, p).  Fixed iteration counts, branch-free selects: the same instruction stream for every input, so threads of a
// warp that invert different values (normalize_batch) do not diverge.  About 5x lower latency than the bit-serial
// binary Euclid above, which matters wherever one inversion sits on a serial path (one per CTA batch of the
// batch-affine rounds, one per final exponentiation).
// in = a*R, out = a^-1 * R; inv(0) = 0.
__device__ __forceinline__ void fpi_mul_small(uint32_t out[13], const uint32_t x[12], uint32_t k) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint64_t t = (uint64_t)x[i] * k + c;
        out[i] = (uint32_t)t;
        c = t >> 32;
    }
    out[12] = (uint32_t)c;
}
// bits [sh, sh + 64) of the 12-limb value x, sh = 32 * w + s chosen by the caller: limbs (x[w+2], x[w+1], x[w]) are
// passed in as (h, m, l)
__device__ __forceinline__ uint64_t fpi_take64(uint32_t h, uint32_t m, uint32_t l, uint32_t s) {
    uint32_t lo = __funnelshift_r(l, m, s), hi = __funnelshift_r(m, h, s);
    return ((uint64_t)hi << 32) | lo;
}
static __device__ __noinline__ Fp fp_inv_pornin(Fp a_mont) {
    uint32_t a[12], b[12], u[12], v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = a_mont.l[i]; b[i] = fp_p_limb(i); u[i] = 0; v[i] = 0; }
    u[0] = 1;
    constexpr uint32_t MINV31 = DG_FP_INV32 & 0x7fffffffu;             // -p^-1 mod 2^31
#pragma unroll 1
    for (int it = 0; it < 25; it++) {
        // n = max(len(a), len(b), 64); abar = (a mod 2^31) + 2^31 * (a >> (n - 33)), same for b
        uint32_t ah = a[2], am = a[1], al = a[0], bh = b[2], bm = b[1], bl = b[0], topw = a[2] | b[2];
        int w = 0;
#pragma unroll
        for (int t = 3; t < 12; t++) {
            bool nz = (a[t] | b[t]) != 0;
            ah = nz ? a[t] : ah; am = nz ? a[t - 1] : am; al = nz ? a[t - 2] : al;
            bh = nz ? b[t] : bh; bm = nz ? b[t - 1] : bm; bl = nz ? b[t - 2] : bl;
            topw = nz ? (a[t] | b[t]) : topw;
            w = nz ? t - 2 : w;
        }
        // (h, m, l) = limbs w+2, w+1, w; the top set bit of a|b sits in limb w+2 unless everything above limb 1 is zero
        uint32_t lz = topw ? __clz(topw) : 32u;                          // n = 32 (w + 3) - lz, or 64 when shorter
        uint32_t s = 32u - lz;                                           // shift inside the (h, m, l) triple: n - 64 - 32 w
        if (w == 0 && topw == 0) s = 0;                                  // both values fit 64 bits: take them exactly
        uint64_t ta = fpi_take64(ah, am, al, s), tb = fpi_take64(bh, bm, bl, s);   // bits [n - 64, n)
        if (s == 32) { ta = ((uint64_t)ah << 32) | am; tb = ((uint64_t)bh << 32) | bm; }   // funnel shift takes s mod 32
        uint64_t abar = (uint64_t)(a[0] & 0x7fffffffu) | ((ta >> 31) << 31);
        uint64_t bbar = (uint64_t)(b[0] & 0x7fffffffu) | ((tb >> 31) << 31);
        long long f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
        for (int j = 0; j < 31; j++) {
            bool odd = abar & 1u, sw = odd && (abar < bbar);
            uint64_t ta2 = sw ? bbar : abar, tb2 = sw ? abar : bbar;
            long long tf = sw ? f1 : f0, tg = sw ? g1 : g0;
            f1 = sw ? f0 : f1; g1 = sw ? g0 : g1;
            f0 = tf; g0 = tg;
            abar = ta2; bbar = tb2;
            abar -= odd ? bbar : 0ull;
            f0 -= odd ? f1 : 0ll;
            g0 -= odd ? g1 : 0ll;
            abar >>= 1;
            f1 <<= 1; g1 <<= 1;
        }
        bool sf0 = f0 < 0, sg0 = g0 < 0, sf1 = f1 < 0, sg1 = g1 < 0;
        uint32_t mf0 = (uint32_t)(sf0 ? -f0 : f0), mg0 = (uint32_t)(sg0 ? -g0 : g0);
        uint32_t mf1 = (uint32_t)(sf1 ? -f1 : f1), mg1 = (uint32_t)(sg1 ? -g1 : g1);
        // (a, b) <- (|f0 a + g0 b|, |f1 a + g1 b|) / 2^31, flipping the signs of a row whose combination is negative
        uint32_t na[12], nb[12];
#pragma unroll
        for (int row = 0; row < 2; row++) {
            uint32_t X[13], Y[13], Z[13];
            fpi_mul_small(X, a, row ? mf1 : mf0);
            fpi_mul_small(Y, b, row ? mg1 : mg0);
            bool sf = row ? sf1 : sf0, sg = row ? sg1 : sg0, neg;
            if (sf == sg) {                                              // same sign: |.| = X + Y
                uint64_t c = 0;
#pragma unroll
                for (int i = 0; i < 13; i++) { c += (uint64_t)X[i] + Y[i]; Z[i] = (uint32_t)c; c >>= 32; }
                neg = sf;
            } else {                                                     // opposite signs: |X - Y|
                uint32_t br = 0;
#pragma unroll
                for (int i = 0; i < 13; i++) {
                    uint64_t d = (uint64_t)X[i] - Y[i] - br;
                    Z[i] = (uint32_t)d;
                    br = (uint32_t)(d >> 63);
                }
                uint32_t cy = br;                                        // negate when X < Y
#pragma unroll
                for (int i = 0; i < 13; i++) {
                    uint64_t t = (uint64_t)(br ? ~Z[i] : Z[i]) + cy;
                    Z[i] = (uint32_t)t;
                    cy = (uint32_t)(t >> 32);
                }
                neg = sf ? !br : (br != 0);
            }
            uint32_t *dst = row ? nb : na;
#pragma unroll
            for (int i = 0; i < 12; i++) dst[i] = (Z[i] >> 31) | (Z[i + 1] << 1);
            if (row) { sf1 ^= neg; sg1 ^= neg; } else { sf0 ^= neg; sg0 ^= neg; }
        }
        // (u, v) <- (f0 u + g0 v, f1 u + g1 v) / 2^31 mod p, operands made non-negative first (x * (-f) = (p - x) * f)
        uint32_t un[12], vn[12], pu[12], pv[12];
        {
            uint32_t br = 0;
#pragma unroll
            for (int i = 0; i < 12; i++) { uint64_t d = (uint64_t)fp_p_limb(i) - u[i] - br; pu[i] = (uint32_t)d; br = (uint32_t)(d >> 63); }
            br = 0;
#pragma unroll
            for (int i = 0; i < 12; i++) { uint64_t d = (uint64_t)fp_p_limb(i) - v[i] - br; pv[i] = (uint32_t)d; br = (uint32_t)(d >> 63); }
        }
#pragma unroll
        for (int row = 0; row < 2; row++) {
            bool sf = row ? sf1 : sf0, sg = row ? sg1 : sg0;
            uint32_t xu[12], xv[12];
#pragma unroll
            for (int i = 0; i < 12; i++) { xu[i] = sf ? pu[i] : u[i]; xv[i] = sg ? pv[i] : v[i]; }
            uint32_t X[13], Y[13], Z[13];
            fpi_mul_small(X, xu, row ? mf1 : mf0);
            fpi_mul_small(Y, xv, row ? mg1 : mg0);
            uint64_t c = 0;
#pragma unroll
            for (int i = 0; i < 13; i++) { c += (uint64_t)X[i] + Y[i]; Z[i] = (uint32_t)c; c >>= 32; }
            uint32_t t = ((Z[0] & 0x7fffffffu) * MINV31) & 0x7fffffffu;
            c = 0;
#pragma unroll
            for (int i = 0; i < 12; i++) { c += (uint64_t)fp_p_limb(i) * t + Z[i]; Z[i] = (uint32_t)c; c >>= 32; }
            Z[12] += (uint32_t)c;
            Fp r;
#pragma unroll
            for (int i = 0; i < 12; i++) r.l[i] = (Z[i] >> 31) | (Z[i + 1] << 1);
            fp_final_sub(r);                                             // < 2p -> [0, p)
            uint32_t *dst = row ? vn : un;
#pragma unroll
            for (int i = 0; i < 12; i++) dst[i] = r.l[i];
        }
#pragma unroll
        for (int i = 0; i < 12; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = un[i]; v[i] = vn[i]; }
    }
    // gcd reached: a = 0, b = 1 and v = (aR)^-1 mod p; for the input 0, b stays p and the answer is 0
    Fp x;
    bool ok = b[0] == 1;
#pragma unroll
    for (int i = 1; i < 12; i++) ok = ok && b[i] == 0;
#pragma unroll
    for (int i = 0; i < 12; i++) x.l[i] = ok ? v[i] : 0u;
    Fp r2;
#pragma unroll
    for (int i = 0; i < 12; i++) r2.l[i] = DGC_R2[i];
    x = fp_mul(x, r2);
    return fp_mul(x, r2);
}

}  // namespace dg
