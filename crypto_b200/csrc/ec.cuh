// Short-Weierstrass (a = 0) group arithmetic for BLS12-381 G1 (F = Fp) and G2 (F = Fp2), written
// once over the generic field interface of fp.cuh / fp2.cuh.
//
//  * XYZZ accumulators (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) for bucket work: mixed add 8M+2S,
//    full add 12M+2S, no inversions; identity is ZZ = 0.
//  * Jacobian (x = X/Z^2, y = Y/Z^3) for double-and-add and as the output format, because that is
//    what ark-ec's Projective<P>{x,y,z} is (the type the reference receives from
//    VariableBaseMSM::msm_bigint / FixedBase::msm / AffineRepr::mul_bigint).
// All formulas are complete in the sense the callers need: identity operands, P + P and P + (-P)
// are detected and handled (the reference's b_g1/b_g2 queries contain identity points,
// legogroth16/src/generator.rs:342,373, and tests feed equal / opposite bases).
#pragma once
#include "fp2.cuh"

namespace dg {

template <class F> struct Affine { F x, y; };            // identity encoded as x = y = 0 (not on the curve)
template <class F> struct XYZZ { F x, y, zz, zzz; };
template <class F> struct Jac { F x, y, z; };

template <class F> __device__ __forceinline__ bool aff_is_inf(const Affine<F> &p) { return fis_zero(p.x) && fis_zero(p.y); }
template <class F> __device__ __forceinline__ bool xyzz_is_inf(const XYZZ<F> &p) { return fis_zero(p.zz); }
template <class F> __device__ __forceinline__ XYZZ<F> xyzz_inf() { return {fzero<F>(), fzero<F>(), fzero<F>(), fzero<F>()}; }
template <class F> __device__ __forceinline__ XYZZ<F> xyzz_from_affine(const Affine<F> &p) {
    bool inf = aff_is_inf(p);
    F one = fone<F>(), z = fzero<F>();
    return {p.x, p.y, fsel(inf, z, one), fsel(inf, z, one)};
}
template <class F> __device__ __forceinline__ XYZZ<F> xyzz_sel(bool c, const XYZZ<F> &a, const XYZZ<F> &b) {
    return {fsel(c, a.x, b.x), fsel(c, a.y, b.y), fsel(c, a.zz, b.zz), fsel(c, a.zzz, b.zzz)};
}

template <class F> __device__ __forceinline__ Affine<F> aff_load(const void *p) {
    const char *q = reinterpret_cast<const char *>(p);
    return {fload<F>(q), fload<F>(q + sizeof(F))};
}
template <class F> __device__ __forceinline__ void aff_store(void *p, const Affine<F> &a) {
    char *q = reinterpret_cast<char *>(p);
    fstore(q, a.x); fstore(q + sizeof(F), a.y);
}
template <class F> __device__ __forceinline__ XYZZ<F> xyzz_load(const void *p) {
    const char *q = reinterpret_cast<const char *>(p);
    return {fload_rw<F>(q), fload_rw<F>(q + sizeof(F)), fload_rw<F>(q + 2 * sizeof(F)), fload_rw<F>(q + 3 * sizeof(F))};
}
template <class F> __device__ __forceinline__ void xyzz_store(void *p, const XYZZ<F> &a) {
    char *q = reinterpret_cast<char *>(p);
    fstore(q, a.x); fstore(q + sizeof(F), a.y); fstore(q + 2 * sizeof(F), a.zz); fstore(q + 3 * sizeof(F), a.zzz);
}
template <class F> __device__ __forceinline__ Jac<F> jac_load(const void *p) {
    const char *q = reinterpret_cast<const char *>(p);
    return {fload_rw<F>(q), fload_rw<F>(q + sizeof(F)), fload_rw<F>(q + 2 * sizeof(F))};
}
template <class F> __device__ __forceinline__ void jac_store(void *p, const Jac<F> &a) {
    char *q = reinterpret_cast<char *>(p);
    fstore(q, a.x); fstore(q + sizeof(F), a.y); fstore(q + 2 * sizeof(F), a.z);
}

// ---- XYZZ ----------------------------------------------------------------------------------
// 2 * (affine point), mdbl-2008-s-1
template <class F> __device__ __noinline__ XYZZ<F> xyzz_dbl_affine(const Affine<F> &p_) {
    const Affine<F> p = p_;                                // local copy: see fp_mul_ni in fp2.cuh
    F u = fdbl(p.y), v = fsqr(u), w = fmul(u, v), s = fmul(p.x, v);
    F xx = fsqr(p.x), m = fadd(fdbl(xx), xx);
    XYZZ<F> r;
    r.x = fsub(fsqr(m), fdbl(s));
    r.y = fsub(fmul(m, fsub(s, r.x)), fmul(w, p.y));
    r.zz = v; r.zzz = w;
    return r;
}
// 2 * (xyzz point), dbl-2008-s-1 (a = 0); caller guarantees p is not the identity
template <class F> __device__ __noinline__ XYZZ<F> xyzz_dbl(const XYZZ<F> &p_) {
    const XYZZ<F> p = p_;
    F u = fdbl(p.y), v = fsqr(u), w = fmul(u, v), s = fmul(p.x, v);
    F xx = fsqr(p.x), m = fadd(fdbl(xx), xx);
    XYZZ<F> r;
    r.x = fsub(fsqr(m), fdbl(s));
    r.y = fsub(fmul(m, fsub(s, r.x)), fmul(w, p.y));
    r.zz = fmul(v, p.zz); r.zzz = fmul(w, p.zzz);
    return r;
}

// acc + q (q affine), madd-2008-s.  Hot loop of bucket accumulation: straight-line for the
// generic case; identity operands are handled by selects, the P == +-Q case by a rare branch.
template <class F> __device__ __forceinline__ XYZZ<F> xyzz_madd(const XYZZ<F> &a, const Affine<F> &q) {
    bool a_inf = xyzz_is_inf(a), q_inf = aff_is_inf(q);
    F u2 = fmul(q.x, a.zz), s2 = fmul(q.y, a.zzz);
    F p = fsub(u2, a.x), r = fsub(s2, a.y);
    XYZZ<F> out;
    if (__builtin_expect(fis_zero(p) && !a_inf && !q_inf, 0)) {
        if (fis_zero(r)) out = xyzz_dbl_affine(q);
        else out = xyzz_inf<F>();
        return out;
    }
    F pp = fsqr(p), ppp = fmul(p, pp), qq = fmul(a.x, pp);
    out.x = fsub(fsub(fsqr(r), ppp), fdbl(qq));
    out.y = fsub(fmul(r, fsub(qq, out.x)), fmul(a.y, ppp));
    out.zz = fmul(a.zz, pp);
    out.zzz = fmul(a.zzz, ppp);
    // identity operands: a = O -> q ; q = O -> a
    XYZZ<F> qx = xyzz_from_affine(q);
    out = xyzz_sel(a_inf, qx, out);
    out = xyzz_sel(q_inf, a, out);
    return out;
}

// a + b, both XYZZ, add-2008-s
template <class F> __device__ __forceinline__ XYZZ<F> xyzz_add(const XYZZ<F> &a, const XYZZ<F> &b) {
    bool a_inf = xyzz_is_inf(a), b_inf = xyzz_is_inf(b);
    F u1 = fmul(a.x, b.zz), u2 = fmul(b.x, a.zz);
    F s1 = fmul(a.y, b.zzz), s2 = fmul(b.y, a.zzz);
    F p = fsub(u2, u1), r = fsub(s2, s1);
    XYZZ<F> out;
    if (__builtin_expect(fis_zero(p) && !a_inf && !b_inf, 0)) {
        if (fis_zero(r)) out = xyzz_dbl(a);
        else out = xyzz_inf<F>();
        return out;
    }
    F pp = fsqr(p), ppp = fmul(p, pp), qq = fmul(u1, pp);
    out.x = fsub(fsub(fsqr(r), ppp), fdbl(qq));
    out.y = fsub(fmul(r, fsub(qq, out.x)), fmul(s1, ppp));
    out.zz = fmul(fmul(a.zz, b.zz), pp);
    out.zzz = fmul(fmul(a.zzz, b.zzz), ppp);
    out = xyzz_sel(a_inf, b, out);
    out = xyzz_sel(b_inf, a, out);
    return out;
}

// XYZZ -> Jacobian without inversion: (X*ZZ, Y*ZZZ, ZZ)   [x = X*ZZ/ZZ^2, y = Y*ZZZ/ZZ^3]
template <class F> __device__ __forceinline__ Jac<F> xyzz_to_jac(const XYZZ<F> &p) {
    if (xyzz_is_inf(p)) return {fone<F>(), fone<F>(), fzero<F>()};     // ark Projective::zero() = (1, 1, 0)
    return {fmul(p.x, p.zz), fmul(p.y, p.zzz), p.zz};
}
template <class F> __device__ __forceinline__ XYZZ<F> jac_to_xyzz(const Jac<F> &p) {
    if (fis_zero(p.z)) return xyzz_inf<F>();
    F zz = fsqr(p.z);
    return {p.x, p.y, zz, fmul(zz, p.z)};
}

// ---- Jacobian ------------------------------------------------------------------------------
template <class F> __device__ __forceinline__ Jac<F> jac_inf() { return {fone<F>(), fone<F>(), fzero<F>()}; }
template <class F> __device__ __forceinline__ bool jac_is_inf(const Jac<F> &p) { return fis_zero(p.z); }

// dbl-2009-l (a = 0): 2M + 5S
template <class F> __device__ __forceinline__ Jac<F> jac_dbl(const Jac<F> &p) {
    F a = fsqr(p.x), b = fsqr(p.y), c = fsqr(b);
    F t = fsub(fsub(fsqr(fadd(p.x, b)), a), c);
    F d = fdbl(t);
    F e = fadd(fdbl(a), a);
    F f = fsqr(e);
    Jac<F> r;
    r.z = fdbl(fmul(p.y, p.z));           // Z = 0 stays 0
    r.x = fsub(f, fdbl(d));
    F c8 = fdbl(fdbl(fdbl(c)));
    r.y = fsub(fmul(e, fsub(d, r.x)), c8);
    return r;
}

// p + q, q affine: madd-2007-bl (7M + 4S)
template <class F> __device__ __forceinline__ Jac<F> jac_madd(const Jac<F> &p, const Affine<F> &q) {
    bool p_inf = jac_is_inf(p), q_inf = aff_is_inf(q);
    F z1z1 = fsqr(p.z);
    F u2 = fmul(q.x, z1z1), s2 = fmul(fmul(q.y, p.z), z1z1);
    F h = fsub(u2, p.x), rr = fsub(s2, p.y);
    if (__builtin_expect(fis_zero(h) && !p_inf && !q_inf, 0)) {
        if (fis_zero(rr)) return jac_dbl(p);
        return jac_inf<F>();
    }
    F hh = fsqr(h), i = fdbl(fdbl(hh)), j = fmul(h, i);
    F r2 = fdbl(rr), v = fmul(p.x, i);
    Jac<F> out;
    out.x = fsub(fsub(fsqr(r2), j), fdbl(v));
    out.y = fsub(fmul(r2, fsub(v, out.x)), fdbl(fmul(p.y, j)));
    out.z = fsub(fsub(fsqr(fadd(p.z, h)), z1z1), hh);
    Jac<F> qj = {q.x, q.y, fone<F>()};
    if (p_inf) out = qj;
    if (q_inf) out = p;
    return out;
}

}  // namespace dg
