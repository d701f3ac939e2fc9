// ark-serialize wire formats of BLS12-381 points on the device ("next" row f4 of SURVEY.md 8f).
//
// Replaces, for vectors of G1 / G2 points, ark-serialize 0.4 CanonicalSerialize::serialize_compressed /
// serialize_uncompressed and CanonicalDeserialize::deserialize_{compressed,uncompressed} with
// Validate::Yes or ::No as the reference reaches them through utils/src/serde_utils.rs:13-33
// (ArkObjectBytes) and the derives on legogroth16/src/data_structures.rs:7-189 (proving / verifying
// keys: loading a key with 2^18..2^19 points per query spends its time in one square root and one
// subgroup check per point).  ark-bls12-381 uses the Zcash / IETF encoding: big-endian coordinates,
// flag bits 0x80 compressed / 0x40 infinity / 0x20 "y is the lexicographically larger root" in
// byte 0, G2 with c1 before c0.
//
// One thread per point.  Decompression: x^3 + b, square root by a^((p+1)/4) (Fp) or the norm
// method (Fp2), root selection by the sort flag.  Subgroup membership by the endomorphism tests
// ark-bls12-381 itself uses (eprint 2021/1130 section 6):  G1  (beta x, y) == -[x^2] P,
// G2  psi(Q) == [x] Q; the oracle checks both against the definition [r] P == O.
//
// Deliberately STRICTER than ark-bls12-381 0.4 on malformed input (arkworks' source is not in /root/reference, so its
// behaviour here is as recalled, not verified): (1) an encoding with the infinity flag must have an all-zero body and
// no sort flag -- arkworks is believed to return the identity without looking at the body; (2) an uncompressed point
// must satisfy the curve equation even with validate == 0 -- arkworks is believed to build it with new_unchecked and
// test the curve and the subgroup only under Validate::Yes.  Every encoding arkworks itself PRODUCES round-trips
// identically; the differences only reject byte strings no serializer emits.  tests/test_wire_formats.py pins both.
#include "common.cuh"
#include "ec.cuh"
#include "fp_inv.cuh"

namespace dg {

enum { SER_OK = 0, SER_MALFORMED = 1, SER_NOT_ON_CURVE = 2, SER_NOT_IN_SUBGROUP = 3 };

// ---- field helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ Fp fp_const(const uint32_t *c) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = c[i];
    return r;
}
static __device__ __noinline__ Fp sfp_mul(Fp a, Fp b) { return fp_mul(a, b); }     // by value: see fp_mul_ni in fp2.cuh
__device__ __forceinline__ Fp fp_to_mont(const Fp &raw) { return sfp_mul(raw, fp_const(DGC_R2)); }
__device__ __forceinline__ Fp fp_from_mont(const Fp &m) {
    Fp one = fp_zero();
    one.l[0] = 1;
    return sfp_mul(m, one);
}
__device__ __forceinline__ bool fp_raw_lt_p(const Fp &a) { return !fp_raw_geq(a, fp_const(DGC_P)); }
// a^e for a 381-bit constant exponent (12 limbs), left-to-right square and multiply
static __device__ __noinline__ Fp fp_pow_const(const Fp &a, const uint32_t *e) {
    Fp r = fp_one();
    for (int bit = 380; bit >= 0; bit--) {
        r = sfp_mul(r, r);
        if ((e[bit >> 5] >> (bit & 31)) & 1u) r = sfp_mul(r, a);
    }
    return r;
}
static __device__ __noinline__ bool fp_sqrt(const Fp &a, Fp &out) {
    out = fp_pow_const(a, DGC_EXP_SQRT);
    return fp_eq(sfp_mul(out, out), a);
}
__device__ __forceinline__ bool fp_lex_largest(const Fp &mont) {          // canonical value > (p - 1) / 2
    Fp c = fp_from_mont(mont);
    return !fp_raw_geq(fp_const(DGC_HALF_P), c);
}
// square root in Fp2 = Fp[u]/(u^2 + 1) by the norm method
static __device__ __noinline__ bool fp2_sqrt(const Fp2 &a, Fp2 &out) {
    if (fp_is_zero(a.c1)) {
        Fp s;
        if (fp_sqrt(a.c0, s)) { out = {s, fp_zero()}; return true; }
        if (fp_sqrt(fp_neg(a.c0), s)) { out = {fp_zero(), s}; return true; }     // (t u)^2 = -t^2
        return false;
    }
    Fp n;
    if (!fp_sqrt(fp_add(sfp_mul(a.c0, a.c0), sfp_mul(a.c1, a.c1)), n)) return false;
    for (int k = 0; k < 2; k++) {
        Fp d = fp_halve(k == 0 ? fp_add(a.c0, n) : fp_sub(a.c0, n));               // halving commutes with the Montgomery factor
        Fp s;
        if (!fp_sqrt(d, s) || fp_is_zero(s)) continue;
        Fp c1 = sfp_mul(a.c1, fp_inv_pornin(fp_dbl(s)));
        out = {s, c1};
        if (feq(fsqr(out), a)) return true;
    }
    return false;
}

// ---- per-curve pieces ---------------------------------------------------------------------------------
__device__ __forceinline__ Fp curve_b(const Fp *) { return fp_const(DGC_B_G1); }
__device__ __forceinline__ Fp2 curve_b(const Fp2 *) { return {fp_const(DGC_B_G1), fp_const(DGC_B_G1)}; }   // 4 (1 + u)
__device__ __forceinline__ bool field_sqrt(const Fp &a, Fp &o) { return fp_sqrt(a, o); }
__device__ __forceinline__ bool field_sqrt(const Fp2 &a, Fp2 &o) { return fp2_sqrt(a, o); }
__device__ __forceinline__ bool lex_largest(const Fp &y) { return fp_lex_largest(y); }
__device__ __forceinline__ bool lex_largest(const Fp2 &y) { return fp_is_zero(y.c1) ? fp_lex_largest(y.c0) : fp_lex_largest(y.c1); }

// [k] P for a 128-bit k, left-to-right double-and-add with mixed additions
template <class F> static __device__ __noinline__ Jac<F> jac_mul_u128(const Affine<F> &p_, uint64_t hi, uint64_t lo) {
    const Affine<F> p = p_;                              // local copy: see fp_mul_ni in fp2.cuh
    Jac<F> acc = jac_inf<F>();
    bool started = false;
    for (int bit = 127; bit >= 0; bit--) {
        bool b = bit >= 64 ? (hi >> (bit - 64)) & 1 : (lo >> bit) & 1;
        if (started) acc = jac_dbl(acc);
        if (b) { acc = jac_madd(acc, p); started = true; }
    }
    return acc;
}
// Jacobian q == affine (x, y)?
template <class F> __device__ __forceinline__ bool jac_eq_affine(const Jac<F> &q, const F &x, const F &y) {
    if (fis_zero(q.z)) return false;
    F zz = fsqr(q.z);
    return feq(q.x, fmul(x, zz)) && feq(q.y, fmul(y, fmul(zz, q.z)));
}
#define DG_X_ABS 0xd201000000010000ull
#define DG_X2_HI 0xac45a4010001a402ull           // x^2 = DG_X2_HI * 2^64 + DG_X2_LO
#define DG_X2_LO 0x100000000ull
static __device__ __noinline__ bool in_subgroup(const Affine<Fp> &p) {     // (beta x, y) == -[x^2] P
    Jac<Fp> q = jac_mul_u128(p, DG_X2_HI, DG_X2_LO);
    return jac_eq_affine(q, sfp_mul(p.x, fp_const(DGC_BETA)), fp_neg(p.y));
}
static __device__ __noinline__ bool in_subgroup(const Affine<Fp2> &p) {    // psi(Q) == [x] Q = -[|x|] Q
    Jac<Fp2> q = jac_mul_u128(p, 0, DG_X_ABS);
    Fp2 px = fmul(fp2_conj(p.x), Fp2{fp_const(DGC_PSI_X0), fp_const(DGC_PSI_X1)});
    Fp2 py = fmul(fp2_conj(p.y), Fp2{fp_const(DGC_PSI_Y0), fp_const(DGC_PSI_Y1)});
    return jac_eq_affine(q, px, fneg(py));
}

// ---- byte order -----------------------------------------------------------------------------------------
// 48 big-endian bytes -> 12 little-endian limbs (plain integer); `src` is 16-byte aligned
__device__ __forceinline__ Fp fp_load_be(const uint8_t *src) {
    const uint4 *q = reinterpret_cast<const uint4 *>(src);
    Fp r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 v = __ldg(q + i);                       // bytes 16 i .. 16 i + 15, most significant first
        r.l[11 - 4 * i] = __byte_perm(v.x, 0, 0x0123);
        r.l[10 - 4 * i] = __byte_perm(v.y, 0, 0x0123);
        r.l[9 - 4 * i] = __byte_perm(v.z, 0, 0x0123);
        r.l[8 - 4 * i] = __byte_perm(v.w, 0, 0x0123);
    }
    return r;
}
__device__ __forceinline__ void fp_store_be(uint8_t *dst, const Fp &a, uint32_t flags) {
    uint4 *q = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 v;
        v.x = __byte_perm(a.l[11 - 4 * i], 0, 0x0123);
        v.y = __byte_perm(a.l[10 - 4 * i], 0, 0x0123);
        v.z = __byte_perm(a.l[9 - 4 * i], 0, 0x0123);
        v.w = __byte_perm(a.l[8 - 4 * i], 0, 0x0123);
        if (i == 0) v.x |= flags;                     // byte 0 of the record is the low byte of v.x
        q[i] = v;
    }
}
// One coordinate of a record (flag bits cleared when it is the first one).  Returns false when it is >= p.
__device__ __forceinline__ bool load_coord(const uint8_t *src, Fp &out, bool has_flags) {
    Fp raw = fp_load_be(src);
    if (has_flags) raw.l[11] &= 0x1fffffffu;
    if (!fp_raw_lt_p(raw)) return false;
    out = fp_to_mont(raw);
    return true;
}
__device__ __forceinline__ bool load_coord(const uint8_t *src, Fp2 &out, bool has_flags) {       // c1 first
    Fp c1 = fp_load_be(src), c0 = fp_load_be(src + 48);
    if (has_flags) c1.l[11] &= 0x1fffffffu;
    if (!fp_raw_lt_p(c1) || !fp_raw_lt_p(c0)) return false;
    out = {fp_to_mont(c0), fp_to_mont(c1)};
    return true;
}
__device__ __forceinline__ void store_coord(uint8_t *dst, const Fp &a, uint32_t flags) { fp_store_be(dst, fp_from_mont(a), flags); }
__device__ __forceinline__ void store_coord(uint8_t *dst, const Fp2 &a, uint32_t flags) {
    fp_store_be(dst, fp_from_mont(a.c1), flags);
    fp_store_be(dst + 48, fp_from_mont(a.c0), 0);
}
__device__ __forceinline__ bool body_is_zero(const uint8_t *src, int bytes) {     // everything but the three flag bits
    const uint32_t *w = reinterpret_cast<const uint32_t *>(src);
    uint32_t acc = __ldg(w) & 0xffffff1fu;
    for (int i = 1; i < bytes / 4; i++) acc |= __ldg(w + i);
    return acc == 0;
}

// ---- kernels --------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_deserialize(const uint8_t *__restrict__ in, uint32_t n, int compressed, int validate,
                                                     Affine<F> *__restrict__ out, uint8_t *__restrict__ status,
                                                     uint32_t *__restrict__ invalid_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int coord = (int)sizeof(F);                        // 48 / 96 encoded bytes per coordinate
    const int rec = compressed ? coord : 2 * coord;
    const uint8_t *src = in + (size_t)i * rec;
    const uint32_t b0 = __ldg(src);
    const bool f_comp = b0 & 0x80, f_inf = b0 & 0x40, f_large = b0 & 0x20;
    Affine<F> p = {fzero<F>(), fzero<F>()};
    int st = SER_OK;
    if (f_comp != (compressed != 0) || (!compressed && f_large)) {
        st = SER_MALFORMED;
    } else if (f_inf) {
        if (f_large || !body_is_zero(src, rec)) st = SER_MALFORMED;
    } else if (!load_coord(src, p.x, true)) {
        st = SER_MALFORMED;
    } else {
        F rhs = fadd(fmul(fsqr(p.x), p.x), curve_b((const F *)nullptr));
        if (compressed) {
            if (!field_sqrt(rhs, p.y)) st = SER_NOT_ON_CURVE;
            else if (lex_largest(p.y) != f_large) p.y = fneg(p.y);
        } else {
            // the y half carries no flag bits: a set top bit makes it >= p and therefore malformed
            F y;
            const bool ok = load_coord(src + coord, y, false);
            if (!ok) st = SER_MALFORMED;
            else if (!feq(fsqr(y), rhs)) st = SER_NOT_ON_CURVE;
            else p.y = y;
        }
        if (st == SER_OK && validate && !in_subgroup(p)) st = SER_NOT_IN_SUBGROUP;
    }
    if (st != SER_OK) {
        p = {fzero<F>(), fzero<F>()};
        atomicAdd(invalid_count, 1u);
    }
    aff_store(&out[i], p);
    status[i] = (uint8_t)st;
}

template <class F>
__global__ void __launch_bounds__(128) k_serialize(const Affine<F> *__restrict__ in, uint32_t n, int compressed, uint8_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int coord = (int)sizeof(F);
    const int rec = compressed ? coord : 2 * coord;
    uint8_t *dst = out + (size_t)i * rec;
    Affine<F> p = aff_load<F>(&in[i]);
    if (aff_is_inf(p)) {
        uint4 *q = reinterpret_cast<uint4 *>(dst);
        for (int k = 0; k < rec / 16; k++) q[k] = make_uint4(k == 0 ? (compressed ? 0xc0u : 0x40u) : 0u, 0u, 0u, 0u);
        return;
    }
    uint32_t flags = 0;
    if (compressed) flags = 0x80u | (lex_largest(p.y) ? 0x20u : 0u);
    store_coord(dst, p.x, flags);
    if (!compressed) store_coord(dst + coord, p.y, 0);
}

// ---- host wrappers --------------------------------------------------------------------------------------
template <class F>
static int32_t deserialize_host(const uint8_t *in, size_t n, int compressed, int validate, uint8_t *out_affine, uint8_t *status,
                                size_t *invalid_count) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (n && (!in || !out_affine)) return fail(DG_ERR_BAD_ARG, "deserialize: null pointer");
    if (n >= (1ull << 31)) return fail(DG_ERR_BAD_ARG, "deserialize: n must be < 2^31");
    if (invalid_count) *invalid_count = 0;
    if (n == 0) return DG_OK;
    const size_t rec = (compressed ? 1 : 2) * sizeof(F), AFF = 2 * sizeof(F);
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(rec * n) + Arena::pad(AFF * n) + Arena::pad(n) + 256, t.stream);
    if (rc) return rc;
    uint8_t *d_in = t.arena.alloc<uint8_t>(rec * n);
    Affine<F> *d_out = t.arena.alloc<Affine<F>>(n);
    uint8_t *d_st = t.arena.alloc<uint8_t>(n);
    uint32_t *d_cnt = t.arena.alloc<uint32_t>(1);
    DG_CUDA(cudaMemcpyAsync(d_in, in, rec * n, cudaMemcpyHostToDevice, t.stream));
    DG_CUDA(cudaMemsetAsync(d_cnt, 0, 4, t.stream));
    DG_LAUNCH(k_deserialize<F>, div_up(n, 128), 128, 0, t.stream, d_in, (uint32_t)n, compressed, validate, d_out, d_st, d_cnt);
    DG_CUDA(cudaMemcpyAsync(out_affine, d_out, AFF * n, cudaMemcpyDeviceToHost, t.stream));
    if (status) DG_CUDA(cudaMemcpyAsync(status, d_st, n, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaMemcpyAsync(t.err_flag_host + 8, d_cnt, 4, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    if (invalid_count) *invalid_count = t.err_flag_host[8];
    return DG_OK;
}

template <class F> static int32_t serialize_host(const uint8_t *affine, size_t n, int compressed, uint8_t *out) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (n && (!affine || !out)) return fail(DG_ERR_BAD_ARG, "serialize: null pointer");
    if (n >= (1ull << 31)) return fail(DG_ERR_BAD_ARG, "serialize: n must be < 2^31");
    if (n == 0) return DG_OK;
    const size_t rec = (compressed ? 1 : 2) * sizeof(F), AFF = 2 * sizeof(F);
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(rec * n) + Arena::pad(AFF * n), t.stream);
    if (rc) return rc;
    Affine<F> *d_in = t.arena.alloc<Affine<F>>(n);
    uint8_t *d_out = t.arena.alloc<uint8_t>(rec * n);
    DG_CUDA(cudaMemcpyAsync(d_in, affine, AFF * n, cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_serialize<F>, div_up(n, 128), 128, 0, t.stream, d_in, (uint32_t)n, compressed, d_out);
    DG_CUDA(cudaMemcpyAsync(out, d_out, rec * n, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

}  // namespace dg

using namespace dg;
extern "C" {
int32_t dg_g1_serialize(const uint8_t *affine, size_t n, int32_t compressed, uint8_t *out) { return serialize_host<Fp>(affine, n, compressed, out); }
int32_t dg_g2_serialize(const uint8_t *affine, size_t n, int32_t compressed, uint8_t *out) { return serialize_host<Fp2>(affine, n, compressed, out); }
int32_t dg_g1_deserialize(const uint8_t *in, size_t n, int32_t compressed, int32_t validate, uint8_t *out_affine, uint8_t *status,
                          size_t *invalid_count) {
    return deserialize_host<Fp>(in, n, compressed, validate, out_affine, status, invalid_count);
}
int32_t dg_g2_deserialize(const uint8_t *in, size_t n, int32_t compressed, int32_t validate, uint8_t *out_affine, uint8_t *status,
                          size_t *invalid_count) {
    return deserialize_host<Fp2>(in, n, compressed, validate, out_affine, status, invalid_count);
}
}
