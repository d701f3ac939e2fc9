// Embarrassingly parallel group operations of the hot path, templated over the base field:
//   fixed-base window tables + batch multiplication   utils::msm::WindowTable (utils/src/msm.rs:8-62)
//                                                     = ark FixedBase::get_window_table / msm
//   independent scalar multiplications                AffineRepr::mul_bigint in cfg_iter! maps
//                                                     (vb_accumulator/src/witness.rs:190,229,278)
//   batch normalisation                               CurveGroup::normalize_batch (witness.rs:193)
//   fused witness update                              witness.rs:269-284
//   fold of per-GPU partial results                   SURVEY.md 8e
#pragma once
#include "common.cuh"
#include "ec.cuh"
#include "fp_inv.cuh"
#include "glv.cuh"
#include "quad.cuh"

namespace dg {

// ---- inversion -------------------------------------------------------------------------------
// a^(p-2) by 4-bit fixed windows (381 squarings + ~95 multiplications); inv(0) = 0.
static __device__ __noinline__ Fp fp_inv(Fp a) {
    Fp tbl[16];
    tbl[0] = fp_one();
    tbl[1] = a;
    for (int i = 2; i < 16; i++) tbl[i] = fp_mul_ni(tbl[i - 1], a);
    // exponent p - 2, 96 nibbles, most significant first
    uint32_t e[12];
#pragma unroll
    for (int i = 0; i < 12; i++) e[i] = fp_p_limb(i);
    e[0] -= 2;                                    // p ends in ...aaab, no borrow
    Fp acc = fp_one();
    for (int nib = 95; nib >= 0; nib--) {
        for (int k = 0; k < 4; k++) acc = fp_mul_ni(acc, acc);
        uint32_t d = (e[nib >> 3] >> ((nib & 7) * 4)) & 15;
        acc = fp_mul_ni(acc, tbl[d]);
    }
    return acc;
}
__device__ __forceinline__ Fp finv(const Fp &a) { return fp_inv_pornin(a); }
__device__ __forceinline__ Fp2 finv(const Fp2 &a) {
    Fp n = fp_inv_pornin(fp_add(fp_mul_ni(a.c0, a.c0), fp_mul_ni(a.c1, a.c1)));
    return {fp_mul_ni(a.c0, n), fp_neg(fp_mul_ni(a.c1, n))};
}

// ---- scalar bit access -------------------------------------------------------------------------
__device__ __forceinline__ void load_scalar(const uint8_t *scalars, size_t i, uint32_t s[9]) {
    const uint4 *sp = reinterpret_cast<const uint4 *>(scalars) + 2 * i;
    uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w; s[8] = 0;
}
__device__ __forceinline__ uint32_t scalar_bits(const uint32_t s[9], int bit, int width) {   // width <= 24
    if (bit >= 256) return 0;
    int word = bit >> 5, off = bit & 31;
    uint64_t two = ((uint64_t)s[word + 1] << 32) | s[word];
    return (uint32_t)(two >> off) & ((1u << width) - 1);
}

// [s]P by signed 4-bit windows with a uniform instruction stream: 64 digits in [-8, 8], MSB first, four doublings and
// ONE addition per digit (always computed, selected away for a zero digit, so the 32 lanes of a warp never diverge):
// 256 doublings + 64 additions + an 8-entry table instead of the 256 + 256 of bit-serial double-and-add.  Table entry
// j holds (j + 1) P as Jacobian plus Z^2 and Z^3, which takes 1M + 1S off every addition (add-2007-bl: 10M + 4S left).
template <class F> struct JacZ { F x, y, z, zz, zzz; };
template <class F> __device__ __forceinline__ JacZ<F> jacz_of(const Jac<F> &p) {
    F zz = fsqr(p.z);
    return {p.x, p.y, p.z, zz, fmul(zz, p.z)};
}
// p + q with q's powers of Z at hand; q is never the identity here (a table entry of a non-identity point)
template <class F> __device__ __forceinline__ Jac<F> jac_add_z(const Jac<F> &p, const JacZ<F> &q) {
    bool p_inf = jac_is_inf(p);
    F z1z1 = fsqr(p.z);
    F u1 = fmul(p.x, q.zz), u2 = fmul(q.x, z1z1);
    F s1 = fmul(p.y, q.zzz), s2 = fmul(fmul(q.y, p.z), z1z1);
    F h = fsub(u2, u1), rr = fsub(s2, s1);
    if (__builtin_expect(fis_zero(h) && !p_inf, 0)) {
        if (fis_zero(rr)) return jac_dbl(p);
        return jac_inf<F>();
    }
    F i = fsqr(fdbl(h)), j = fmul(h, i);
    F r2 = fdbl(rr), v = fmul(u1, i);
    Jac<F> out;
    out.x = fsub(fsub(fsqr(r2), j), fdbl(v));
    out.y = fsub(fmul(r2, fsub(v, out.x)), fdbl(fmul(s1, j)));
    out.z = fmul(fsub(fsub(fsqr(fadd(p.z, q.z)), z1z1), q.zz), h);
    if (p_inf) out = {q.x, q.y, q.z};
    return out;
}
// signed 4-bit digits of v[0 .. words), least significant first: d = nibble + carry, d > 8 -> d - 16 and carry; the
// carry out of the top nibble is returned (always 0 when that nibble is <= 7).  mag nibble = 8 | (|d| - 1), or 0 for d == 0.
template <int WORDS> __device__ __forceinline__ uint32_t recode_w4(const uint32_t *v, uint32_t mag[WORDS], uint64_t &neg) {
    neg = 0;
    uint32_t carry = 0;
#pragma unroll
    for (int w = 0; w < WORDS; w++) {
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t d = ((v[w] >> (4 * k)) & 15u) + carry;
            bool ng = d > 8;
            carry = ng ? 1u : 0u;
            uint32_t a = ng ? 16u - d : d;              // 0 .. 8
            m |= (a ? (8u | (a - 1)) : 0u) << (4 * k);
            neg |= (uint64_t)(ng ? 1u : 0u) << (8 * w + k);
        }
        mag[w] = m;
    }
    return carry;                                       // a 65th digit (0 or 1) for 256-bit integers whose top nibble exceeds 7
}
template <class F> __device__ __forceinline__ void scalar_mul_table(const Affine<F> &p, JacZ<F> tbl[8]) {
    Jac<F> p1 = {p.x, p.y, fone<F>()};
    Jac<F> p2 = jac_dbl(p1), p3 = jac_madd(p2, p), p4 = jac_dbl(p2);
    Jac<F> p5 = jac_madd(p4, p), p6 = jac_dbl(p3), p7 = jac_madd(p6, p), p8 = jac_dbl(p4);
    tbl[0] = {p.x, p.y, fone<F>(), fone<F>(), fone<F>()};
    tbl[1] = jacz_of(p2); tbl[2] = jacz_of(p3); tbl[3] = jacz_of(p4);
    tbl[4] = jacz_of(p5); tbl[5] = jacz_of(p6); tbl[6] = jacz_of(p7); tbl[7] = jacz_of(p8);
}
// out-of-line group operations: one copy of each in the kernel instead of one per call site
// (operands copied into locals first: see fp_mul_ni in fp2.cuh)
template <class F> static __device__ __noinline__ Jac<F> jac_dbl_ni(const Jac<F> &p_) { Jac<F> p = p_; return jac_dbl(p); }
template <class F> static __device__ __noinline__ Jac<F> jac_add_z_ni(const Jac<F> &p_, const JacZ<F> &q_) { Jac<F> p = p_; JacZ<F> q = q_; return jac_add_z(p, q); }
template <class F> static __device__ __noinline__ void scalar_mul_table_ni(const Affine<F> &p_, JacZ<F> *tbl) { Affine<F> p = p_; scalar_mul_table(p, tbl); }

// [s] P (+ [s2] V) for 256-bit integers, signed 4-bit windows, uniform instruction stream.
//   canonical scalars (< r), points in the prime-order subgroup -- what every reference call site passes
//   (mul_bigint(fr.into_bigint()) on deserialised, hence validated, points) -- and glv != 0:
//       GLV split s = +-k1 - k2 lambda; the 127-bit halves of every scalar share ONE chain of 128 doublings (Shamir's
//       trick), phi is applied to the table entry on the fly (one multiplication by beta): 128 dbl + 64 add per scalar;
//   otherwise (s >= r is not a field element but a legal BigInt): the plain 65-digit chain, 256 dbl + 64 add.
// vtbl != nullptr adds [s2] V with V's eight multiples read from global memory (k_w4_table): the fused accumulator
// witness update d * C_i + v * V (vb_accumulator/src/witness.rs:269-284) on one doubling chain.
template <class F> struct W4Digits { uint32_t mag[4][8]; uint64_t ng[4]; bool sgn[4]; uint32_t top[4]; };
template <class F> __device__ __forceinline__ bool w4_split(const uint32_t s_in[9], uint32_t *mag1, uint64_t &ng1, bool &sgn1, uint32_t *mag2,
                                                           uint64_t &ng2, bool &sgn2) {
    constexpr uint32_t RM[8] = {DG_R0, DG_R1, DG_R2, DG_R3, DG_R4, DG_R5, DG_R6, DG_R7};
    uint32_t s[8], t[8], borrow = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {                      // t = r - s
        uint64_t d = (uint64_t)RM[k] - s_in[k] - borrow;
        t[k] = (uint32_t)d;
        borrow = (uint32_t)(d >> 63);
    }
    bool t_zero = true, decided = false, t_less = false;
#pragma unroll
    for (int k = 7; k >= 0; k--) {
        t_zero &= t[k] == 0;
        if (!decided && t[k] != s_in[k]) { t_less = t[k] < s_in[k]; decided = true; }
    }
    if (borrow || t_zero) return false;                // s >= r
    const bool flip = t_less;
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = flip ? t[k] : s_in[k];
    uint32_t k1[4], k2[4];
    bool neg1;
    glv_split(s, k1, neg1, k2);
    recode_w4<4>(k1, mag1, ng1);
    recode_w4<4>(k2, mag2, ng2);
    sgn1 = flip != neg1;                               // [s] P = sigma ( +-[k1] P - [k2] phi(P) )
    sgn2 = !flip;
    return true;
}
template <class F>
static __device__ __noinline__ Jac<F> scalar_mul_w4(const Affine<F> &p, const uint32_t s_in[9], int glv, const JacZ<F> *vtbl = nullptr,
                                                    const uint32_t *s2_in = nullptr) {
    W4Digits<F> dg_;
    int ndig = 65, halves = 1;
    bool split = false;
    const bool p_inf = aff_is_inf(p);
    if (vtbl && fis_zero(vtbl[0].z)) vtbl = nullptr;       // V is the identity (k_w4_table marks it with Z = 0)
    if (glv) {
        split = w4_split<F>(s_in, dg_.mag[0], dg_.ng[0], dg_.sgn[0], dg_.mag[1], dg_.ng[1], dg_.sgn[1]);
        if (split && vtbl) split = w4_split<F>(s2_in, dg_.mag[2], dg_.ng[2], dg_.sgn[2], dg_.mag[3], dg_.ng[3], dg_.sgn[3]);
        if (split) { ndig = 32; halves = vtbl ? 4 : 2; }
    }
    if (!split) {
        dg_.top[0] = recode_w4<8>(s_in, dg_.mag[0], dg_.ng[0]);
        dg_.sgn[0] = false;
        if (vtbl) {                                        // second scalar on the same 65-digit chain: halves 0 (P) and 2 (V)
            dg_.top[2] = recode_w4<8>(s2_in, dg_.mag[2], dg_.ng[2]);
            dg_.sgn[2] = false;
        }
    }
    JacZ<F> tbl[8];
    if (!p_inf) scalar_mul_table_ni(p, tbl);
    Fp beta;
#pragma unroll
    for (int k = 0; k < 12; k++) beta.l[k] = sizeof(F) > 48 ? DGC_GLV_BETA_G2[k] : DGC_GLV_BETA_G1[k];
    Jac<F> acc = jac_inf<F>();
#pragma unroll 1
    for (int dig = ndig - 1; dig >= 0; dig--) {
#pragma unroll 1
        for (int k = 0; k < 4; k++) acc = jac_dbl_ni(acc);
#pragma unroll 1
        for (int half = 0; half < 4; half++) {
            const bool use = split ? half < halves : (half == 0 || (half == 2 && vtbl != nullptr));
            if (!use) continue;                                                  // uniform across the warp's split / unsplit lanes only
            const bool on_v = half >= 2;
            if (!on_v && p_inf) continue;
            uint32_t nib = dig < 64 ? (dg_.mag[half][dig >> 3] >> (4 * (dig & 7))) & 15u : (dg_.top[half] ? 8u : 0u);
            bool nz = (nib & 8u) != 0, neg = dig < 64 && ((((dg_.ng[half] >> dig) & 1u) != 0) != dg_.sgn[half]);
            JacZ<F> q = on_v ? vtbl[nib & 7u] : tbl[nib & 7u];
            if (split && (half & 1)) q.x = glv_mul_beta(q.x, beta);             // phi((X, Y, Z)) = (beta X, Y, Z)
            q.y = fcneg(q.y, neg);
            Jac<F> sum = jac_add_z_ni(acc, q);
            acc.x = fsel(nz, sum.x, acc.x); acc.y = fsel(nz, sum.y, acc.y); acc.z = fsel(nz, sum.z, acc.z);
        }
    }
    return acc;
}
// the eight multiples of V with their Z^2, Z^3 (table of the shared point of the fused update), one thread
template <class F> __global__ void k_w4_table(const Affine<F> *v, JacZ<F> *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Affine<F> p = aff_load<F>(v);
    JacZ<F> tbl[8];
    if (aff_is_inf(p)) {
        for (int j = 0; j < 8; j++) tbl[j] = {fone<F>(), fone<F>(), fzero<F>(), fzero<F>(), fzero<F>()};
    } else {
        scalar_mul_table(p, tbl);
    }
    for (int j = 0; j < 8; j++) out[j] = tbl[j];
}

template <class F>
__global__ void __launch_bounds__(128) k_batch_mul(const Affine<F> *__restrict__ points, const uint8_t *__restrict__ scalars,
                                                   uint32_t m, Jac<F> *__restrict__ out, int glv, int one_scalar = 0,
                                                   const Affine<F> *__restrict__ add = nullptr) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t s[9];
    load_scalar(scalars, one_scalar ? 0 : i, s);           // one_scalar: every point is multiplied by scalars[0]
    Affine<F> p = aff_load<F>(&points[i]);
    Jac<F> r = scalar_mul_w4(p, s, glv);
    if (add) r = jac_madd(r, aff_load<F>(&add[i]));        // + add[i]: the "compress" step of the SnarkPack GIPA rounds
    jac_store(&out[i], r);
}

// ---- fixed-base tables -------------------------------------------------------------------------
// g_outer[k] = 2^(k*window) * g   (sequential doubling chain, one thread)
template <class F>
__global__ void k_fixed_outer(const Affine<F> *g, int window, int outerc, Jac<F> *gouter) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Affine<F> p = aff_load<F>(g);
    Jac<F> cur = aff_is_inf(p) ? jac_inf<F>() : Jac<F>{p.x, p.y, fone<F>()};
    for (int k = 0; k < outerc; k++) {
        jac_store(&gouter[k], cur);
        for (int t = 0; t < window; t++) cur = jac_dbl(cur);
    }
}
// the same chain by one quad (quad.cuh): 3 product waves per doubling instead of 7-9 dependent multiplications by one thread
// (255 doublings: ~0.85 ms instead of ~1.7 ms in G1, ~1 ms instead of ~5 ms in G2)
template <class F>
__global__ void __launch_bounds__(32) k_fixed_outer_quad(const Affine<F> *g, int window, int outerc, Jac<F> *gouter) {
    extern __shared__ __align__(16) unsigned char dg_smem_bq[];
    QuadWS<F> &ws = reinterpret_cast<QuadWS<F> *>(dg_smem_bq)[0];
    const QuadCtx qc = quad_ctx<QuadWide<F>::value>();
    if (!qc.active || qc.qi != 0 || blockIdx.x != 0) return;
    enum { ACC = DG_Q_ACC };
    const Affine<F> p = aff_load<F>(g);
    const bool inf = aff_is_inf(p);
    F one = fone<F>();
    ws.v[4 * ACC + qc.ql] = inf ? fzero<F>() : (qc.ql == 0 ? p.x : qc.ql == 1 ? p.y : one);
    __syncwarp(qc.mask);
    for (int k = 0; k < outerc; k++) {
        if (qc.ql == 0 && qc.part == 0) {
            XYZZ<F> r = {ws.v[4 * ACC], ws.v[4 * ACC + 1], ws.v[4 * ACC + 2], ws.v[4 * ACC + 3]};
            jac_store(&gouter[k], xyzz_to_jac(r));
        }
        __syncwarp(qc.mask);
        if (!inf && k + 1 < outerc)
            for (int t = 0; t < window; t++) quad_dbl(ws, ACC, ACC, qc);
    }
}
// Jacobian + Jacobian via XYZZ (used off the hot path only)
template <class F> __device__ __forceinline__ Jac<F> jac_add_slow(const Jac<F> &a, const Jac<F> &b) {
    return xyzz_to_jac(xyzz_add(jac_to_xyzz(a), jac_to_xyzz(b)));
}
// table[k][j] = j * g_outer[k] as Jacobian (normalised afterwards); thread per (k, j)
template <class F>
__global__ void __launch_bounds__(128) k_fixed_rows(const Jac<F> *__restrict__ gouter, int window, int outerc,
                                                    Jac<F> *__restrict__ rows) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t in_window = 1u << window;
    if (tid >= (uint32_t)outerc * in_window) return;
    uint32_t k = tid >> window, j = tid & (in_window - 1);
    Jac<F> base = jac_load<F>(&gouter[k]);
    XYZZ<F> b = jac_to_xyzz(base), acc = xyzz_inf<F>();
    // ark's last row only has 2^(255 - (outerc-1)*window) entries; the rest stays the identity
    if (k == (uint32_t)outerc - 1 && j >= (1u << (255 - (outerc - 1) * window))) j = 0;
    for (int bit = window - 1; bit >= 0; bit--) {
        if (!xyzz_is_inf(acc)) acc = xyzz_dbl(acc);
        if ((j >> bit) & 1) acc = xyzz_add(acc, b);
    }
    jac_store(&rows[tid], xyzz_to_jac(acc));
}

// FixedBase::windowed_mul for every scalar: sum_k table[k][bits_k(s)]  (mixed XYZZ additions)
template <class F>
__global__ void __launch_bounds__(128) k_fixed_mul_many(const Affine<F> *__restrict__ table, int window, int outerc,
                                                        const uint8_t *__restrict__ scalars, uint32_t m,
                                                        Jac<F> *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t s[9];
    load_scalar(scalars, i, s);
    XYZZ<F> acc = xyzz_inf<F>();
    for (int k = 0; k < outerc; k++) {
        uint32_t idx = scalar_bits(s, k * window, window);
        Affine<F> q = aff_load<F>(&table[((size_t)k << window) + idx]);
        acc = xyzz_madd(acc, q);
    }
    jac_store(&out[i], xyzz_to_jac(acc));
}

// out[i] = [a_i] P_i + [b_i] V  (V through its window table), Jacobian
template <class F>
__global__ void __launch_bounds__(128) k_batch_mul_add_fixed(const Affine<F> *__restrict__ points, const uint8_t *__restrict__ sa,
                                                             const Affine<F> *__restrict__ table, int window, int outerc,
                                                             const uint8_t *__restrict__ sb, uint32_t m, Jac<F> *__restrict__ out, int glv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t s[9];
    load_scalar(sa, i, s);
    Affine<F> p = aff_load<F>(&points[i]);
    XYZZ<F> acc = jac_to_xyzz(scalar_mul_w4(p, s, glv));
    load_scalar(sb, i, s);
    for (int k = 0; k < outerc; k++) {
        uint32_t idx = scalar_bits(s, k * window, window);
        Affine<F> q = aff_load<F>(&table[((size_t)k << window) + idx]);
        acc = xyzz_madd(acc, q);
    }
    jac_store(&out[i], xyzz_to_jac(acc));
}

// out[i] = [a_i] P_i + [b_i] V on ONE doubling chain per element (no window table): the latency-optimal form of the fused
// update for batches that cannot fill the GPU (10^4 witnesses = 79 CTAs), where the 255 sequential doublings of building
// V's window table would cost more than the table saves.
template <class F>
__global__ void __launch_bounds__(128) k_batch_mul_add_joint(const Affine<F> *__restrict__ points, const uint8_t *__restrict__ sa,
                                                             const JacZ<F> *__restrict__ vtbl, const uint8_t *__restrict__ sb, uint32_t m,
                                                             Jac<F> *__restrict__ out, int glv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t s[9], s2[9];
    load_scalar(sa, i, s);
    load_scalar(sb, i, s2);
    Affine<F> p = aff_load<F>(&points[i]);
    jac_store(&out[i], scalar_mul_w4(p, s, glv, vtbl, s2));
}
// Same result with TWO threads per element, for batches too small to fill the GPU with one thread each: warps 0-1 of a
// CTA compute [a_i] P_i, warps 2-3 [b_i] V for the same 64 elements (each its own 128-doubling chain, 64 additions),
// and the halves meet in shared memory for one final addition.  The chain per thread is a third shorter than the
// joint one (128 dbl + 64 add instead of 128 + 128), which is what a latency-bound launch pays for.
template <class F>
__global__ void __launch_bounds__(128) k_batch_mul_add_split(const Affine<F> *__restrict__ points, const uint8_t *__restrict__ sa,
                                                             const JacZ<F> *__restrict__ vtbl, const uint8_t *__restrict__ sb, uint32_t m,
                                                             Jac<F> *__restrict__ out, int glv) {
    __shared__ Jac<F> half[64];
    const uint32_t role = threadIdx.x >> 6, slot = threadIdx.x & 63;          // role is warp-uniform
    const uint32_t i = blockIdx.x * 64 + slot;
    Jac<F> r = jac_inf<F>();
    if (i < m) {
        uint32_t s[9];
        load_scalar(role ? sb : sa, i, s);
        if (role == 0) {
            Affine<F> p = aff_load<F>(&points[i]);
            r = scalar_mul_w4(p, s, glv);
        } else {
            Affine<F> none = {fzero<F>(), fzero<F>()};                          // identity: only the V halves run
            r = scalar_mul_w4(none, s, glv, vtbl, s);
        }
    }
    if (role == 1) half[slot] = r;
    __syncthreads();
    if (role == 0 && i < m) jac_store(&out[i], jac_add_slow(r, half[slot]));
}

// ---- quad-cooperative scalar multiplication (small batches) ------------------------------------------------------------
// A batch too small to fill the GPU with one thread per element (10^4 accumulator witnesses = 80 threads per SM) is bound by
// the LATENCY of one element's chain: 128 doublings + 64 additions of ~7-11 dependent field multiplications each, ~1 us per
// multiplication for a lone thread.  Here a quad (4 lanes, 12 over Fp2: quad.cuh) owns one chain and issues every point
// operation as 3-4 waves of independent products, ~2x shorter per operation.  Per chain: the eight multiples j P and
// their images phi(j P) = (beta x, y) go to a global scratch table as XYZZ records, then the signed 4-bit digits of the GLV
// halves (or the 65-digit chain of an integer >= r) are walked with quad_dbl / quad_add.  The fused update
// [a_i] P_i + [b_i] V runs the two products of an element on two quads and a second small kernel adds the pairs.
template <class F> struct BatchQuadGeom {
    static constexpr int QP = sizeof(F) > 48 ? 16 : 32;                                   // quads per CTA
    static constexpr int THREADS = QuadLanes<QuadWide<F>::value>::cta_threads(QP);
};
// chain q: v == nullptr: [sa_q] points[q] -> out[q];  v != nullptr: even q: [sa_(q/2)] points[q/2], odd q: [sb_(q/2)] V -> out[q]
template <class F>
__global__ void __launch_bounds__(BatchQuadGeom<F>::THREADS) k_batch_mul_quad(const Affine<F> *__restrict__ points, const uint8_t *__restrict__ sa,
                                                                               const Affine<F> *__restrict__ v, const uint8_t *__restrict__ sb,
                                                                               uint32_t nchains, int glv, XYZZ<F> *__restrict__ tables,
                                                                               XYZZ<F> *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char dg_smem_bq[];
    QuadWS<F> *wsall = reinterpret_cast<QuadWS<F> *>(dg_smem_bq);
    const QuadCtx qc = quad_ctx<QuadWide<F>::value>();
    if (!qc.active) return;
    const uint32_t q = blockIdx.x * BatchQuadGeom<F>::QP + qc.qi;
    if (q >= nchains) return;                                                              // quad-uniform; no CTA barriers below
    QuadWS<F> &ws = wsall[qc.qi];
    enum { ACC = DG_Q_ACC, ITEM = DG_Q_ITEM };
    const uint32_t elem = v ? q >> 1 : q;
    const bool on_v = v != nullptr && (q & 1u);
    const Affine<F> *base = on_v ? v : &points[elem];
    uint32_t s[9];
    load_scalar(on_v ? sb : sa, elem, s);
    // digits (every lane of the quad computes the same values)
    W4Digits<F> dg_;
    int ndig = 65;
    bool split = false;
    if (glv) split = w4_split<F>(s, dg_.mag[0], dg_.ng[0], dg_.sgn[0], dg_.mag[1], dg_.ng[1], dg_.sgn[1]);
    if (split) ndig = 32;
    else { dg_.top[0] = recode_w4<8>(s, dg_.mag[0], dg_.ng[0]); dg_.sgn[0] = false; }
    // table: T[j] = (j + 1) P, T[8 + j] = phi(T[j]), j < 8
    XYZZ<F> *T = tables + (size_t)16 * q;
    Fp beta;
#pragma unroll
    for (int k = 0; k < 12; k++) beta.l[k] = sizeof(F) > 48 ? DGC_GLV_BETA_G2[k] : DGC_GLV_BETA_G1[k];
    const Affine<F> p = aff_load<F>(base);
    const bool p_inf = aff_is_inf(p);
    if (!p_inf) {
        // ITEM = P as XYZZ (x, y, 1, 1); ACC walks P, 2P, ..., 8P
        F one = fone<F>();
        ws.v[4 * ITEM + qc.ql] = qc.ql == 0 ? p.x : qc.ql == 1 ? p.y : one;
        ws.v[4 * ACC + qc.ql] = qc.ql == 0 ? p.x : qc.ql == 1 ? p.y : one;
        __syncwarp(qc.mask);
        for (int j = 0; j < 8; j++) {
            if (j == 1) quad_dbl(ws, ACC, ACC, qc);
            else if (j > 1) quad_add(ws, ACC, ACC, ITEM, qc);
            F c = ws.v[4 * ACC + qc.ql];
            if (qc.part == 0) {
                fstore(reinterpret_cast<char *>(&T[j]) + qc.ql * sizeof(F), c);
                if (split) fstore(reinterpret_cast<char *>(&T[8 + j]) + qc.ql * sizeof(F), qc.ql == 0 ? glv_mul_beta(c, beta) : c);
            }
            __syncwarp(qc.mask);
        }
    }
    __syncwarp(qc.mask);
    quad_set_inf(ws, ACC, qc);
    if (!p_inf) {
#pragma unroll 1
        for (int dig = ndig - 1; dig >= 0; dig--) {
#pragma unroll 1
            for (int k = 0; k < 4; k++)
                if (!fis_zero(ws.v[4 * ACC + 2])) quad_dbl(ws, ACC, ACC, qc);
#pragma unroll 1
            for (int half = 0; half < (split ? 2 : 1); half++) {
                uint32_t nib = dig < 64 ? (dg_.mag[half][dig >> 3] >> (4 * (dig & 7))) & 15u : (dg_.top[half] ? 8u : 0u);
                if (!(nib & 8u)) continue;                                                 // quad-uniform
                const bool neg = dig < 64 && ((((dg_.ng[half] >> dig) & 1u) != 0) != dg_.sgn[half]);
                const XYZZ<F> *e = &T[(half ? 8 : 0) + (nib & 7u)];
                F c = fload_rw<F>(reinterpret_cast<const char *>(e) + qc.ql * sizeof(F));
                if (qc.ql == 1) c = fcneg(c, neg);
                ws.v[4 * ITEM + qc.ql] = c;
                __syncwarp(qc.mask);
                quad_add(ws, ACC, ACC, ITEM, qc);
            }
        }
    }
    quad_store(ws, ACC, &out[q], qc);
}
// out[i] = in[2i] + in[2i + 1] (pairs == 1) or in[i] (pairs == 0), as Jacobian (ark Projective layout)
template <class F>
__global__ void __launch_bounds__(BatchQuadGeom<F>::THREADS) k_quad_finish(const XYZZ<F> *__restrict__ in, uint32_t m, int pairs, Jac<F> *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char dg_smem_bq[];
    QuadWS<F> *wsall = reinterpret_cast<QuadWS<F> *>(dg_smem_bq);
    const QuadCtx qc = quad_ctx<QuadWide<F>::value>();
    if (!qc.active) return;
    const uint32_t i = blockIdx.x * BatchQuadGeom<F>::QP + qc.qi;
    if (i >= m) return;
    QuadWS<F> &ws = wsall[qc.qi];
    enum { ACC = DG_Q_ACC, ITEM = DG_Q_ITEM };
    quad_load(ws, ACC, &in[pairs ? 2 * i : i], qc);
    if (pairs) {
        quad_load(ws, ITEM, &in[2 * i + 1], qc);
        quad_add(ws, ACC, ACC, ITEM, qc);
    }
    __syncwarp(qc.mask);
    if (qc.ql == 0 && qc.part == 0) {
        XYZZ<F> r = {ws.v[4 * ACC], ws.v[4 * ACC + 1], ws.v[4 * ACC + 2], ws.v[4 * ACC + 3]};
        jac_store(&out[i], xyzz_to_jac(r));
    }
}

// ---- normalize_batch ---------------------------------------------------------------------------
// Montgomery's trick on chunks of DG_NORM_CHUNK points per thread: one inversion per chunk.
#define DG_NORM_CHUNK 16
template <class F>
__global__ void __launch_bounds__(128) k_normalize(const Jac<F> *__restrict__ in, uint32_t m, Affine<F> *__restrict__ out,
                                                   F *__restrict__ prefix) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo = t * DG_NORM_CHUNK;
    if (lo >= m) return;
    uint32_t hi = lo + DG_NORM_CHUNK < m ? lo + DG_NORM_CHUNK : m;
    F acc = fone<F>();
    for (uint32_t i = lo; i < hi; i++) {
        F z = fload_rw<F>(&in[i].z);
        fstore(&prefix[i], acc);
        if (!fis_zero(z)) acc = fmul(acc, z);
    }
    F inv = finv(acc);
    for (uint32_t i = hi; i-- > lo;) {
        Jac<F> p = jac_load<F>(&in[i]);
        Affine<F> a;
        if (fis_zero(p.z)) {
            a.x = fzero<F>(); a.y = fzero<F>();
        } else {
            F zi = fmul(inv, fload_rw<F>(&prefix[i]));
            inv = fmul(inv, p.z);
            F zi2 = fsqr(zi);
            a.x = fmul(p.x, zi2);
            a.y = fmul(p.y, fmul(zi2, zi));
        }
        aff_store(&out[i], a);
    }
}

// ---- precomputed multiples for resident bases ------------------------------------------------
// row k (k >= 1) of the table is 2^(c*k) * P_i; thread i walks its own doubling chain.
template <class F>
__global__ void __launch_bounds__(128) k_precompute_rows(const Affine<F> *__restrict__ bases, uint32_t n, int c, int rows,
                                                         Jac<F> *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = aff_load<F>(&bases[i]);
    Jac<F> cur = aff_is_inf(p) ? jac_inf<F>() : Jac<F>{p.x, p.y, fone<F>()};
    for (int k = 1; k < rows; k++) {
        for (int t = 0; t < c; t++) cur = jac_dbl(cur);
        jac_store(&out[(size_t)(k - 1) * n + i], cur);
    }
}

// ---- fold ----------------------------------------------------------------------------------------
template <class F> __global__ void k_fold_jac(const Jac<F> *in, uint32_t k, Jac<F> *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<F> acc = xyzz_inf<F>();
    for (uint32_t i = 0; i < k; i++) acc = xyzz_add(acc, jac_to_xyzz(jac_load<F>(&in[i])));
    jac_store(out, xyzz_to_jac(acc));
}

// Same fold with one pointer per partial: on a multi-GPU context the pointers are the other devices' result
// buffers, loaded over NVLink peer access by this one thread ("all-reduce under the group law").
template <class F> __global__ void k_fold_jac_ptrs(PtrList in, uint32_t k, Jac<F> *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<F> acc = xyzz_inf<F>();
    for (uint32_t i = 0; i < k; i++) acc = xyzz_add(acc, jac_to_xyzz(jac_load<F>(in.p[i])));
    jac_store(out, xyzz_to_jac(acc));
}

}  // namespace dg
