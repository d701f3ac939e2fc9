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

namespace dg {

// ---- inversion -------------------------------------------------------------------------------
// a^(p-2) by 4-bit fixed windows (381 squarings + ~95 multiplications); inv(0) = 0.
static __device__ __noinline__ Fp fp_inv(const Fp &a) {
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
__device__ __forceinline__ Fp finv(const Fp &a) { return fp_inv_binary(a); }
__device__ __forceinline__ Fp2 finv(const Fp2 &a) {
    Fp n = fp_inv_binary(fp_add(fp_mul_ni(a.c0, a.c0), fp_mul_ni(a.c1, a.c1)));
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

// [s]P, MSB-first double-and-add with a uniform instruction stream (the addition is always
// computed and selected by the bit, so the 32 lanes of a warp never diverge).
template <class F> __device__ __forceinline__ Jac<F> scalar_mul(const Affine<F> &p, const uint32_t s[9]) {
    Jac<F> acc = jac_inf<F>();
    int top = 255;
    for (int bit = top; bit >= 0; bit--) {
        acc = jac_dbl(acc);
        Jac<F> sum = jac_madd(acc, p);
        bool b = (s[bit >> 5] >> (bit & 31)) & 1;
        acc.x = fsel(b, sum.x, acc.x); acc.y = fsel(b, sum.y, acc.y); acc.z = fsel(b, sum.z, acc.z);
    }
    return acc;
}

template <class F>
__global__ void __launch_bounds__(128) k_batch_mul(const Affine<F> *__restrict__ points, const uint8_t *__restrict__ scalars,
                                                   uint32_t m, Jac<F> *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t s[9];
    load_scalar(scalars, i, s);
    Affine<F> p = aff_load<F>(&points[i]);
    jac_store(&out[i], scalar_mul(p, s));
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
                                                             const uint8_t *__restrict__ sb, uint32_t m, Jac<F> *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t s[9];
    load_scalar(sa, i, s);
    Affine<F> p = aff_load<F>(&points[i]);
    XYZZ<F> acc = jac_to_xyzz(scalar_mul(p, s));
    load_scalar(sb, i, s);
    for (int k = 0; k < outerc; k++) {
        uint32_t idx = scalar_bits(s, k * window, window);
        Affine<F> q = aff_load<F>(&table[((size_t)k << window) + idx]);
        acc = xyzz_madd(acc, q);
    }
    jac_store(&out[i], xyzz_to_jac(acc));
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
