/* CPU oracle for the BLS12-381 hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (crypto_b200/, libdockgpu.so) never does.
 *
 * What it restates: the arkworks 0.4 algorithms the reference calls on its hot path
 * (SURVEY.md section 8a / Appendix B).  The arithmetic is in the third-party crates
 * ark-ec ^0.4.1, ark-ff ^0.4.1, ark-bls12-381 ^0.4.0 (/root/reference/Cargo.toml:32-46),
 * which are NOT vendored and whose exact patch version is unpinned (no Cargo.lock), so
 * the published algorithms are restated:
 *   ref_msm_g1/g2          VariableBaseMSM::msm_bigint (msm_bigint_wnaf): window rule
 *                          c = ln_without_floats(n)+2, signed digits, rayon-over-windows
 *                          -> OpenMP over windows.  Call sites: legogroth16/src/prover.rs:286,
 *                          299,363,592; bbs_plus/src/setup.rs:145; schnorr_pok/src/
 *                          pok_generalized_pedersen.rs:97; vb_accumulator/src/witness.rs:415
 *   ref_fixed_base_*       FixedBase::get_window_table / msm behind utils/src/msm.rs:18-62
 *   ref_batch_mul_g1       AffineRepr::mul_bigint per element (vb_accumulator/src/witness.rs:190)
 *   ref_normalize_batch_g1 CurveGroup::normalize_batch (vb_accumulator/src/witness.rs:193)
 *   ref_multi_miller_loop, ref_final_exp
 *                          ark_ec::models::bls12 (G2Prepared line coefficients, M-twist ell,
 *                          conjugate for x<0; final exponentiation chain of eprint 2020/875)
 *                          reached from utils/src/randomized_pairing_check.rs:204-214,
 *                          bbs_plus/src/proof.rs:494, legogroth16/src/verifier.rs:69-80
 *
 * PARITY STATUS: "parity unpinned" by the reference (it holds no golden vectors for this
 * path).  Pinned instead against oracle/bls12_381.py (big-int, two independent pairing
 * derivations) by tests/test_oracle.py and the vectors in tests/golden/.
 *
 * 6 x 64-bit Montgomery limbs (R = 2^384), unsigned __int128 products: the same
 * representation as ark-ff's MontBackend, so Montgomery bytes cross unchanged.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "bls_consts.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fp_t;
typedef struct { fp_t c0, c1; } fp2_t;
typedef struct { fp2_t c0, c1, c2; } fp6_t;
typedef struct { fp6_t c0, c1; } fp12_t;

/* ------------------------------------------------------------------ Fp --- */
static inline void fp_set_zero(fp_t *r) { memset(r, 0, sizeof *r); }
static inline void fp_set_one(fp_t *r) { memcpy(r->l, FPC_R_ONE, 48); }
static inline int fp_is_zero(const fp_t *a) { uint64_t t = 0; for (int i = 0; i < 6; i++) t |= a->l[i]; return t == 0; }
static inline int fp_eq(const fp_t *a, const fp_t *b) { return memcmp(a, b, 48) == 0; }

static inline int fp_geq_p(const uint64_t a[6]) {
    for (int i = 5; i >= 0; i--) { if (a[i] > FPC_P[i]) return 1; if (a[i] < FPC_P[i]) return 0; }
    return 1;
}
static inline void fp_sub_p(uint64_t a[6]) {
    u128 br = 0;
    for (int i = 0; i < 6; i++) { u128 t = (u128)a[i] - FPC_P[i] - br; a[i] = (uint64_t)t; br = (t >> 64) & 1; }
}
static inline void fp_add(fp_t *r, const fp_t *a, const fp_t *b) {
    u128 c = 0; uint64_t t[6];
    for (int i = 0; i < 6; i++) { c += (u128)a->l[i] + b->l[i]; t[i] = (uint64_t)c; c >>= 64; }
    if (fp_geq_p(t)) fp_sub_p(t);
    memcpy(r->l, t, 48);
}
static inline void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) {
    u128 br = 0; uint64_t t[6];
    for (int i = 0; i < 6; i++) { u128 d = (u128)a->l[i] - b->l[i] - br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 6; i++) { c += (u128)t[i] + FPC_P[i]; t[i] = (uint64_t)c; c >>= 64; } }
    memcpy(r->l, t, 48);
}
static inline void fp_neg(fp_t *r, const fp_t *a) { fp_t z; fp_set_zero(&z); if (fp_is_zero(a)) { *r = z; return; } fp_sub(r, &z, a); }

/* CIOS Montgomery multiplication */
static inline void fp_mul(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[8] = {0};
    for (int i = 0; i < 6; i++) {
        u128 c = 0;
        for (int j = 0; j < 6; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[6]; t[6] = (uint64_t)c; t[7] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FP_INV64;
        c = (u128)m * FPC_P[0] + t[0]; c >>= 64;
        for (int j = 1; j < 6; j++) { c += (u128)m * FPC_P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[6]; t[5] = (uint64_t)c; t[6] = t[7] + (uint64_t)(c >> 64);
    }
    if (t[6] || fp_geq_p(t)) fp_sub_p(t);
    memcpy(r->l, t, 48);
}
static inline void fp_sqr(fp_t *r, const fp_t *a) { fp_mul(r, a, a); }

static void fp_pow(fp_t *r, const fp_t *a, const uint64_t *e, int nlimbs) {
    fp_t acc; fp_set_one(&acc);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        fp_sqr(&acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void fp_inv(fp_t *r, const fp_t *a) {           /* a^(p-2); inv(0) = 0 */
    uint64_t e[6]; memcpy(e, FPC_P, 48); e[0] -= 2;
    fp_pow(r, a, e, 6);
}

/* ------------------------------------------------------------------ Fp2 -- */
static inline void fp2_set_zero(fp2_t *r) { fp_set_zero(&r->c0); fp_set_zero(&r->c1); }
static inline void fp2_set_one(fp2_t *r) { fp_set_one(&r->c0); fp_set_zero(&r->c1); }
static inline int fp2_is_zero(const fp2_t *a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static inline int fp2_eq(const fp2_t *a, const fp2_t *b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static inline void fp2_add(fp2_t *r, const fp2_t *a, const fp2_t *b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static inline void fp2_sub(fp2_t *r, const fp2_t *a, const fp2_t *b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static inline void fp2_neg(fp2_t *r, const fp2_t *a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static inline void fp2_conj(fp2_t *r, const fp2_t *a) { r->c0 = a->c0; fp_neg(&r->c1, &a->c1); }
static inline void fp2_mul(fp2_t *r, const fp2_t *a, const fp2_t *b) {
    fp_t t0, t1, s0, s1, m;
    fp_mul(&t0, &a->c0, &b->c0); fp_mul(&t1, &a->c1, &b->c1);
    fp_add(&s0, &a->c0, &a->c1); fp_add(&s1, &b->c0, &b->c1);
    fp_mul(&m, &s0, &s1);
    fp_sub(&r->c0, &t0, &t1);
    fp_sub(&m, &m, &t0); fp_sub(&r->c1, &m, &t1);
}
static inline void fp2_sqr(fp2_t *r, const fp2_t *a) {
    fp_t s, d, m;
    fp_add(&s, &a->c0, &a->c1); fp_sub(&d, &a->c0, &a->c1);
    fp_mul(&m, &a->c0, &a->c1);
    fp_mul(&r->c0, &s, &d); fp_add(&r->c1, &m, &m);
}
static inline void fp2_mul_fp(fp2_t *r, const fp2_t *a, const fp_t *k) { fp_mul(&r->c0, &a->c0, k); fp_mul(&r->c1, &a->c1, k); }
static inline void fp2_mul_xi(fp2_t *r, const fp2_t *a) {       /* times (1+u) */
    fp_t t0, t1; fp_sub(&t0, &a->c0, &a->c1); fp_add(&t1, &a->c0, &a->c1); r->c0 = t0; r->c1 = t1;
}
static void fp2_inv(fp2_t *r, const fp2_t *a) {
    fp_t n, t; fp_sqr(&n, &a->c0); fp_sqr(&t, &a->c1); fp_add(&n, &n, &t); fp_inv(&n, &n);
    fp_mul(&r->c0, &a->c0, &n); fp_mul(&t, &a->c1, &n); fp_neg(&r->c1, &t);
}

/* ------------------------------------------------------------------ Fp6 -- */
static void fp6_add(fp6_t *r, const fp6_t *a, const fp6_t *b) { fp2_add(&r->c0, &a->c0, &b->c0); fp2_add(&r->c1, &a->c1, &b->c1); fp2_add(&r->c2, &a->c2, &b->c2); }
static void fp6_sub(fp6_t *r, const fp6_t *a, const fp6_t *b) { fp2_sub(&r->c0, &a->c0, &b->c0); fp2_sub(&r->c1, &a->c1, &b->c1); fp2_sub(&r->c2, &a->c2, &b->c2); }
static void fp6_neg(fp6_t *r, const fp6_t *a) { fp2_neg(&r->c0, &a->c0); fp2_neg(&r->c1, &a->c1); fp2_neg(&r->c2, &a->c2); }
static void fp6_mul(fp6_t *r, const fp6_t *a, const fp6_t *b) {
    fp2_t t0, t1, t2, x, y, c0, c1, c2;
    fp2_mul(&t0, &a->c0, &b->c0); fp2_mul(&t1, &a->c1, &b->c1); fp2_mul(&t2, &a->c2, &b->c2);
    fp2_add(&x, &a->c1, &a->c2); fp2_add(&y, &b->c1, &b->c2); fp2_mul(&c0, &x, &y);
    fp2_sub(&c0, &c0, &t1); fp2_sub(&c0, &c0, &t2); fp2_mul_xi(&c0, &c0); fp2_add(&c0, &c0, &t0);
    fp2_add(&x, &a->c0, &a->c1); fp2_add(&y, &b->c0, &b->c1); fp2_mul(&c1, &x, &y);
    fp2_sub(&c1, &c1, &t0); fp2_sub(&c1, &c1, &t1); fp2_mul_xi(&x, &t2); fp2_add(&c1, &c1, &x);
    fp2_add(&x, &a->c0, &a->c2); fp2_add(&y, &b->c0, &b->c2); fp2_mul(&c2, &x, &y);
    fp2_sub(&c2, &c2, &t0); fp2_sub(&c2, &c2, &t2); fp2_add(&c2, &c2, &t1);
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static void fp6_mul_v(fp6_t *r, const fp6_t *a) { fp2_t t; fp2_mul_xi(&t, &a->c2); r->c2 = a->c1; r->c1 = a->c0; r->c0 = t; }
static void fp6_inv(fp6_t *r, const fp6_t *a) {
    fp2_t t0, t1, t2, x, d;
    fp2_sqr(&t0, &a->c0); fp2_mul(&x, &a->c1, &a->c2); fp2_mul_xi(&x, &x); fp2_sub(&t0, &t0, &x);
    fp2_sqr(&t1, &a->c2); fp2_mul_xi(&t1, &t1); fp2_mul(&x, &a->c0, &a->c1); fp2_sub(&t1, &t1, &x);
    fp2_sqr(&t2, &a->c1); fp2_mul(&x, &a->c0, &a->c2); fp2_sub(&t2, &t2, &x);
    fp2_mul(&d, &a->c2, &t1); fp2_mul(&x, &a->c1, &t2); fp2_add(&d, &d, &x); fp2_mul_xi(&d, &d);
    fp2_mul(&x, &a->c0, &t0); fp2_add(&d, &d, &x);
    fp2_inv(&d, &d);
    fp2_mul(&r->c0, &t0, &d); fp2_mul(&r->c1, &t1, &d); fp2_mul(&r->c2, &t2, &d);
}

/* ------------------------------------------------------------------ Fp12 - */
static void fp12_set_one(fp12_t *r) { memset(r, 0, sizeof *r); fp_set_one(&r->c0.c0.c0); }
static int fp12_is_zero(const fp12_t *a) { const uint64_t *w = (const uint64_t *)a; uint64_t t = 0; for (int i = 0; i < 72; i++) t |= w[i]; return t == 0; }
static void fp12_mul(fp12_t *r, const fp12_t *a, const fp12_t *b) {
    fp6_t t0, t1, x, y, c1;
    fp6_mul(&t0, &a->c0, &b->c0); fp6_mul(&t1, &a->c1, &b->c1);
    fp6_add(&x, &a->c0, &a->c1); fp6_add(&y, &b->c0, &b->c1); fp6_mul(&c1, &x, &y);
    fp6_sub(&c1, &c1, &t0); fp6_sub(&c1, &c1, &t1);
    fp6_mul_v(&x, &t1); fp6_add(&r->c0, &t0, &x);
    r->c1 = c1;
}
static void fp12_sqr(fp12_t *r, const fp12_t *a) { fp12_mul(r, a, a); }
static void fp12_conj(fp12_t *r, const fp12_t *a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); }
static void fp12_inv(fp12_t *r, const fp12_t *a) {
    fp6_t d, t;
    fp6_mul(&d, &a->c0, &a->c0); fp6_mul(&t, &a->c1, &a->c1); fp6_mul_v(&t, &t); fp6_sub(&d, &d, &t);
    fp6_inv(&d, &d);
    fp6_mul(&r->c0, &a->c0, &d); fp6_mul(&t, &a->c1, &d); fp6_neg(&r->c1, &t);
}
static void fp12_frobenius(fp12_t *r, const fp12_t *a, int k) {
    const uint64_t (*g)[2][6] = k == 1 ? FPC_FROB1 : k == 2 ? FPC_FROB2 : FPC_FROB3;
    fp2_t *dst[6] = {&r->c0.c0, &r->c1.c0, &r->c0.c1, &r->c1.c1, &r->c0.c2, &r->c1.c2};
    const fp2_t *src[6] = {&a->c0.c0, &a->c1.c0, &a->c0.c1, &a->c1.c1, &a->c0.c2, &a->c1.c2};
    for (int i = 0; i < 6; i++) {            /* basis element w^i picks up xi^(i(p^k-1)/6) */
        fp2_t t = *src[i], c;
        if (k & 1) fp2_conj(&t, &t);
        memcpy(c.c0.l, g[i][0], 48); memcpy(c.c1.l, g[i][1], 48);
        fp2_mul(dst[i], &t, &c);
    }
}
/* f *= (c0 + c1 v + c4 v w) -- ark Fp12::mul_by_014 */
static void fp12_mul_by_014(fp12_t *f, const fp2_t *c0, const fp2_t *c1, const fp2_t *c4) {
    fp12_t s; memset(&s, 0, sizeof s);
    s.c0.c0 = *c0; s.c0.c1 = *c1; s.c1.c1 = *c4;
    fp12_mul(f, f, &s);
}
#define BLS_X_ABS 0xd201000000010000ULL
static void fp12_exp_by_x(fp12_t *r, const fp12_t *a) {   /* a^|x| then conjugate (x<0) */
    fp12_t acc; fp12_set_one(&acc);
    for (int i = 63; i >= 0; i--) {
        fp12_sqr(&acc, &acc);
        if ((BLS_X_ABS >> i) & 1) fp12_mul(&acc, &acc, a);
    }
    fp12_conj(r, &acc);
}

/* ------------------------------------------------------ scalars / windows -- */
static int ceil_log2(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }
static int ln_without_floats(size_t n) { return ceil_log2(n) * 69 / 100; }
static int msm_window_size(size_t n) { return n < 32 ? 3 : ln_without_floats(n) + 2; }
static int fixed_base_window_size(size_t n) { return n < 32 ? 3 : ln_without_floats(n); }

/* ark make_digits: signed radix-2^c digits of a 256-bit canonical integer */
static void make_digits(int32_t *out, const uint64_t s[4], int c, int nd) {
    int64_t radix = (int64_t)1 << c, half = radix >> 1, carry = 0;
    for (int i = 0; i < nd; i++) {
        int bit = i * c;
        uint64_t v = 0;
        if (bit < 256) {
            int w = bit >> 6, off = bit & 63;
            v = s[w] >> off;
            if (off + c > 64 && w + 1 < 4) v |= s[w + 1] << (64 - off);
            v &= (uint64_t)radix - 1;
        }
        int64_t d = (int64_t)v + carry;
        carry = (d + half) >> c;
        d -= carry << c;
        if (i == nd - 1) d += carry << c;
        out[i] = (int32_t)d;
    }
}

/* -------------------------------------------------------- G1 / G2 --------- */
#define FE fp_t
#define FN(x) fp_##x
#define PN(x) g1_##x
#include "ec_tmpl.h"
#undef FE
#undef FN
#undef PN
#define FE fp2_t
#define FN(x) fp2_##x
#define PN(x) g2_##x
#include "ec_tmpl.h"
#undef FE
#undef FN
#undef PN

/* packed wire records of the C ABI: identity = all-zero coordinates */
static void g1_load(g1_aff *p, const uint8_t *b) {
    memcpy(&p->x, b, 48); memcpy(&p->y, b + 48, 48);
    p->inf = fp_is_zero(&p->x) && fp_is_zero(&p->y);
}
static void g1_store(uint8_t *b, const g1_aff *p) {
    if (p->inf) { memset(b, 0, 96); return; }
    memcpy(b, &p->x, 48); memcpy(b + 48, &p->y, 48);
}
static void g2_load(g2_aff *p, const uint8_t *b) {
    memcpy(&p->x, b, 96); memcpy(&p->y, b + 96, 96);
    p->inf = fp2_is_zero(&p->x) && fp2_is_zero(&p->y);
}
static void g2_store(uint8_t *b, const g2_aff *p) {
    if (p->inf) { memset(b, 0, 192); return; }
    memcpy(b, &p->x, 96); memcpy(b + 96, &p->y, 96);
}

/* -------------------------------------------------------- pairing --------- */
typedef struct { fp2_t c0, c1, c2; } ell_coeff;
#define N_ELL 68

static void g2_prepare(ell_coeff *co, const g2_aff *q) {
    fp2_t rx = q->x, ry = q->y, rz; fp2_set_one(&rz);
    fp_t two_inv; memcpy(two_inv.l, FPC_TWO_INV, 48);
    fp2_t B; memcpy(B.c0.l, FPC_B_G1, 48); memcpy(B.c1.l, FPC_B_G1, 48);   /* 4 + 4u */
    int n = 0;
    for (int i = 62; i >= 0; i--) {
        {   /* doubling step */
            fp2_t a, b, c, e, f, g, h, ii, j, e2, t;
            fp2_mul(&a, &rx, &ry); fp2_mul_fp(&a, &a, &two_inv);
            fp2_sqr(&b, &ry); fp2_sqr(&c, &rz);
            fp2_add(&t, &c, &c); fp2_add(&t, &t, &c); fp2_mul(&e, &B, &t);
            fp2_add(&f, &e, &e); fp2_add(&f, &f, &e);
            fp2_add(&g, &b, &f); fp2_mul_fp(&g, &g, &two_inv);
            fp2_add(&h, &ry, &rz); fp2_sqr(&h, &h); fp2_add(&t, &b, &c); fp2_sub(&h, &h, &t);
            fp2_sub(&ii, &e, &b);
            fp2_sqr(&j, &rx);
            fp2_sqr(&e2, &e);
            fp2_sub(&t, &b, &f); fp2_mul(&rx, &a, &t);
            fp2_sqr(&ry, &g); fp2_add(&t, &e2, &e2); fp2_add(&t, &t, &e2); fp2_sub(&ry, &ry, &t);
            fp2_mul(&rz, &b, &h);
            co[n].c0 = ii; fp2_add(&t, &j, &j); fp2_add(&co[n].c1, &t, &j); fp2_neg(&co[n].c2, &h);
            n++;
        }
        if ((BLS_X_ABS >> i) & 1) {   /* addition step */
            fp2_t theta, lam, c, d, e, f, g, h, j, t;
            fp2_mul(&t, &q->y, &rz); fp2_sub(&theta, &ry, &t);
            fp2_mul(&t, &q->x, &rz); fp2_sub(&lam, &rx, &t);
            fp2_sqr(&c, &theta); fp2_sqr(&d, &lam);
            fp2_mul(&e, &lam, &d); fp2_mul(&f, &rz, &c); fp2_mul(&g, &rx, &d);
            fp2_add(&h, &e, &f); fp2_sub(&h, &h, &g); fp2_sub(&h, &h, &g);
            fp2_mul(&rx, &lam, &h);
            fp2_sub(&t, &g, &h); fp2_mul(&t, &theta, &t); fp2_mul(&ry, &e, &ry); fp2_sub(&ry, &t, &ry);
            fp2_mul(&rz, &rz, &e);
            fp2_mul(&j, &theta, &q->x); fp2_mul(&t, &lam, &q->y); fp2_sub(&j, &j, &t);
            co[n].c0 = j; fp2_neg(&co[n].c1, &theta); co[n].c2 = lam;
            n++;
        }
    }
}

static void ell(fp12_t *f, const ell_coeff *co, const g1_aff *p) {
    fp2_t c1, c2;
    fp2_mul_fp(&c2, &co->c2, &p->y);
    fp2_mul_fp(&c1, &co->c1, &p->x);
    fp12_mul_by_014(f, &co->c0, &c1, &c2);
}

/* ================================================================= exports == */
#define EXPORT __attribute__((visibility("default")))

EXPORT int ref_msm_window_size(size_t n) { return msm_window_size(n); }
EXPORT int ref_fixed_base_window_size(size_t n) { return fixed_base_window_size(n); }

/* out: Jacobian 144 B (ark Projective x,y,z) */
EXPORT void ref_msm_g1(const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac) {
    g1_aff *b = (g1_aff *)malloc(sizeof(g1_aff) * (n ? n : 1));
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) g1_load(&b[i], bases + 96 * i);
    g1_jac r; g1_msm_bigint(&r, b, (const uint64_t *)scalars, n);
    memcpy(out_jac, &r, 144);
    free(b);
}
EXPORT void ref_msm_g2(const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac) {
    g2_aff *b = (g2_aff *)malloc(sizeof(g2_aff) * (n ? n : 1));
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) g2_load(&b[i], bases + 192 * i);
    g2_jac r; g2_msm_bigint(&r, b, (const uint64_t *)scalars, n);
    memcpy(out_jac, &r, 288);
    free(b);
}

EXPORT void ref_normalize_batch_g1(const uint8_t *jac, size_t n, uint8_t *out_aff) {
    g1_aff *a = (g1_aff *)malloc(sizeof(g1_aff) * (n ? n : 1));
    g1_normalize_batch(a, (const g1_jac *)jac, n);
    for (size_t i = 0; i < n; i++) g1_store(out_aff + 96 * i, &a[i]);
    free(a);
}
EXPORT void ref_normalize_batch_g2(const uint8_t *jac, size_t n, uint8_t *out_aff) {
    g2_aff *a = (g2_aff *)malloc(sizeof(g2_aff) * (n ? n : 1));
    g2_normalize_batch(a, (const g2_jac *)jac, n);
    for (size_t i = 0; i < n; i++) g2_store(out_aff + 192 * i, &a[i]);
    free(a);
}

/* [s_i] P_i, out Jacobian m x 144 B */
EXPORT void ref_batch_mul_g1(const uint8_t *points, const uint8_t *scalars, size_t m, uint8_t *out_jac) {
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < m; i++) {
        g1_aff p; g1_load(&p, points + 96 * i);
        g1_jac r; g1_mul_bigint(&r, &p, (const uint64_t *)(scalars + 32 * i));
        memcpy(out_jac + 144 * i, &r, 144);
    }
}
EXPORT void ref_batch_mul_g2(const uint8_t *points, const uint8_t *scalars, size_t m, uint8_t *out_jac) {
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < m; i++) {
        g2_aff p; g2_load(&p, points + 192 * i);
        g2_jac r; g2_mul_bigint(&r, &p, (const uint64_t *)(scalars + 32 * i));
        memcpy(out_jac + 288 * i, &r, 288);
    }
}

/* WindowTable::new(hint_n, g) + multiply_many(scalars): out Jacobian m x 144 B.
 * Also returns the table geometry (window, num_windows) like utils/src/msm.rs:8-13. */
EXPORT void ref_fixed_base_mul_many_g1(const uint8_t *point, size_t hint_n, const uint8_t *scalars, size_t m,
                                       uint8_t *out_jac, int *window_out, int *num_windows_out) {
    g1_aff g; g1_load(&g, point);
    int window = fixed_base_window_size(hint_n), outerc;
    g1_aff *table = g1_fixed_base_table(&g, window, &outerc);
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < m; i++) {
        g1_jac r; g1_windowed_mul(&r, table, window, outerc, (const uint64_t *)(scalars + 32 * i));
        memcpy(out_jac + 144 * i, &r, 144);
    }
    if (window_out) *window_out = window;
    if (num_windows_out) *num_windows_out = outerc;
    free(table);
}
EXPORT void ref_fixed_base_mul_many_g2(const uint8_t *point, size_t hint_n, const uint8_t *scalars, size_t m,
                                       uint8_t *out_jac, int *window_out, int *num_windows_out) {
    g2_aff g; g2_load(&g, point);
    int window = fixed_base_window_size(hint_n), outerc;
    g2_aff *table = g2_fixed_base_table(&g, window, &outerc);
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < m; i++) {
        g2_jac r; g2_windowed_mul(&r, table, window, outerc, (const uint64_t *)(scalars + 32 * i));
        memcpy(out_jac + 288 * i, &r, 288);
    }
    if (window_out) *window_out = window;
    if (num_windows_out) *num_windows_out = outerc;
    free(table);
}
/* Row-major window table itself (num_windows x 2^window affine records, 96 B each). */
EXPORT void ref_fixed_base_table_g1(const uint8_t *point, int window, uint8_t *out_table) {
    g1_aff g; g1_load(&g, point);
    int outerc; g1_aff *table = g1_fixed_base_table(&g, window, &outerc);
    size_t tot = (size_t)outerc << window;
    for (size_t i = 0; i < tot; i++) g1_store(out_table + 96 * i, &table[i]);
    free(table);
}

/* sum_i a_i * b_i over the integers (no reduction) as a 576-bit little-endian integer: the O(n) side of the
 * known-discrete-log identity  sum s_i (k_i G) = (sum s_i k_i mod r) G  that tests and bench.py use to check
 * MSMs far larger than the oracle's own Pippenger could finish in seconds.  a, b: n x 32 B canonical LE. */
EXPORT void ref_scalar_dot_wide(const uint8_t *a, const uint8_t *b, size_t n, uint8_t out[72]) {
    uint64_t acc[9] = {0};
    for (size_t i = 0; i < n; i++) {
        const uint64_t *x = (const uint64_t *)(a + 32 * i), *y = (const uint64_t *)(b + 32 * i);
        uint64_t prod[8] = {0};
        for (int j = 0; j < 4; j++) {
            unsigned __int128 carry = 0;
            for (int k = 0; k < 4; k++) {
                unsigned __int128 t = (unsigned __int128)x[j] * y[k] + prod[j + k] + carry;
                prod[j + k] = (uint64_t)t;
                carry = t >> 64;
            }
            prod[j + 4] = (uint64_t)carry;
        }
        unsigned __int128 c = 0;
        for (int k = 0; k < 8; k++) {
            c += (unsigned __int128)acc[k] + prod[k];
            acc[k] = (uint64_t)c;
            c >>= 64;
        }
        acc[8] += (uint64_t)c;
    }
    memcpy(out, acc, 72);
}

/* k_i * G1 generator for synthetic bases (fixed-base, fast): out affine m x 96 B */
EXPORT void ref_g1_generator_muls(const uint8_t *scalars, size_t m, uint8_t *out_aff) {
    g1_aff g; memcpy(&g.x, FPC_G1_X, 48); memcpy(&g.y, FPC_G1_Y, 48); g.inf = 0;
    int window = 12, outerc;
    g1_aff *table = g1_fixed_base_table(&g, window, &outerc);
    g1_jac *r = (g1_jac *)malloc(sizeof(g1_jac) * (m ? m : 1));
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < m; i++) g1_windowed_mul(&r[i], table, window, outerc, (const uint64_t *)(scalars + 32 * i));
    g1_aff *a = (g1_aff *)malloc(sizeof(g1_aff) * (m ? m : 1));
    const size_t CH = 4096;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t c0 = 0; c0 < m; c0 += CH) g1_normalize_batch(a + c0, r + c0, (m - c0 < CH) ? m - c0 : CH);
    for (size_t i = 0; i < m; i++) g1_store(out_aff + 96 * i, &a[i]);
    free(a); free(r); free(table);
}
EXPORT void ref_g2_generator_muls(const uint8_t *scalars, size_t m, uint8_t *out_aff) {
    g2_aff g; memcpy(&g.x.c0, FPC_G2_X0, 48); memcpy(&g.x.c1, FPC_G2_X1, 48);
    memcpy(&g.y.c0, FPC_G2_Y0, 48); memcpy(&g.y.c1, FPC_G2_Y1, 48); g.inf = 0;
    int window = 10, outerc;
    g2_aff *table = g2_fixed_base_table(&g, window, &outerc);
    g2_jac *r = (g2_jac *)malloc(sizeof(g2_jac) * (m ? m : 1));
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < m; i++) g2_windowed_mul(&r[i], table, window, outerc, (const uint64_t *)(scalars + 32 * i));
    g2_aff *a = (g2_aff *)malloc(sizeof(g2_aff) * (m ? m : 1));
    const size_t CH = 4096;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t c0 = 0; c0 < m; c0 += CH) g2_normalize_batch(a + c0, r + c0, (m - c0 < CH) ? m - c0 : CH);
    for (size_t i = 0; i < m; i++) g2_store(out_aff + 192 * i, &a[i]);
    free(a); free(r); free(table);
}

/* Bls12::multi_miller_loop: identity pairs dropped; chunks of 4 pairs in parallel;
 * out = Fp12 576 B (c0.c0.c0 ... c1.c2.c1, Montgomery) */
EXPORT void ref_multi_miller_loop(const uint8_t *g1s, const uint8_t *g2s, size_t k, uint8_t *out_fp12) {
    g1_aff *ps = (g1_aff *)malloc(sizeof(g1_aff) * (k ? k : 1));
    ell_coeff *cos = (ell_coeff *)malloc(sizeof(ell_coeff) * N_ELL * (k ? k : 1));
    size_t m = 0;
    for (size_t i = 0; i < k; i++) {
        g1_aff p; g2_aff q; g1_load(&p, g1s + 96 * i); g2_load(&q, g2s + 192 * i);
        if (p.inf || q.inf) continue;
        ps[m] = p; g2_prepare(cos + N_ELL * m, &q); m++;
    }
    size_t nchunks = (m + 3) / 4;
    fp12_t *part = (fp12_t *)malloc(sizeof(fp12_t) * (nchunks ? nchunks : 1));
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t ch = 0; ch < nchunks; ch++) {
        size_t lo = ch * 4, hi = lo + 4 < m ? lo + 4 : m;
        fp12_t f; fp12_set_one(&f);
        int idx = 0;
        for (int i = 62; i >= 0; i--) {
            fp12_sqr(&f, &f);
            for (size_t j = lo; j < hi; j++) ell(&f, &cos[N_ELL * j + idx], &ps[j]);
            idx++;
            if ((BLS_X_ABS >> i) & 1) {
                for (size_t j = lo; j < hi; j++) ell(&f, &cos[N_ELL * j + idx], &ps[j]);
                idx++;
            }
        }
        part[ch] = f;
    }
    fp12_t f; fp12_set_one(&f);
    for (size_t ch = 0; ch < nchunks; ch++) fp12_mul(&f, &f, &part[ch]);
    fp12_conj(&f, &f);
    memcpy(out_fp12, &f, 576);
    free(part); free(cos); free(ps);
}

/* Bls12::final_exponentiation: returns 0 (None) iff input is zero */
EXPORT int ref_final_exp(const uint8_t *in_fp12, uint8_t *out_fp12) {
    fp12_t f; memcpy(&f, in_fp12, 576);
    if (fp12_is_zero(&f)) return 0;
    fp12_t f1, f2, r, y0, y1, y2;
    fp12_conj(&f1, &f);
    fp12_inv(&f2, &f);
    fp12_mul(&r, &f1, &f2);
    f2 = r;
    fp12_frobenius(&r, &r, 2);
    fp12_mul(&r, &r, &f2);
    fp12_sqr(&y0, &r);
    fp12_exp_by_x(&y1, &r);
    fp12_conj(&y2, &r);
    fp12_mul(&y1, &y1, &y2);
    fp12_exp_by_x(&y2, &y1);
    fp12_conj(&y1, &y1);
    fp12_mul(&y1, &y1, &y2);
    fp12_exp_by_x(&y2, &y1);
    fp12_frobenius(&y1, &y1, 1);
    fp12_mul(&y1, &y1, &y2);
    fp12_mul(&r, &r, &y0);
    fp12_exp_by_x(&y0, &y1);
    fp12_exp_by_x(&y2, &y0);
    fp12_frobenius(&y0, &y1, 2);
    fp12_conj(&y1, &y1);
    fp12_mul(&y1, &y1, &y2);
    fp12_mul(&y1, &y1, &y0);
    fp12_mul(&r, &r, &y1);
    memcpy(out_fp12, &r, 576);
    return 1;
}

EXPORT void ref_fp12_mul(const uint8_t *a, const uint8_t *b, uint8_t *out) {
    fp12_t x, y, r; memcpy(&x, a, 576); memcpy(&y, b, 576); fp12_mul(&r, &x, &y); memcpy(out, &r, 576);
}
/* GT exponentiation by a canonical 256-bit integer (PairingOutput::mul_bigint) */
EXPORT void ref_fp12_pow(const uint8_t *a, const uint8_t *scalar, uint8_t *out) {
    fp12_t x, acc; memcpy(&x, a, 576); fp12_set_one(&acc);
    const uint64_t *s = (const uint64_t *)scalar;
    for (int i = 255; i >= 0; i--) { fp12_sqr(&acc, &acc); if ((s[i >> 6] >> (i & 63)) & 1) fp12_mul(&acc, &acc, &x); }
    memcpy(out, &acc, 576);
}
EXPORT void ref_fp12_one(uint8_t *out) { fp12_t o; fp12_set_one(&o); memcpy(out, &o, 576); }

/* Fp helpers for tests (Montgomery in/out) */
EXPORT void ref_fp_mul(const uint8_t *a, const uint8_t *b, uint8_t *out) {
    fp_t x, y, r; memcpy(&x, a, 48); memcpy(&y, b, 48); fp_mul(&r, &x, &y); memcpy(out, &r, 48);
}
EXPORT void ref_fp_inv(const uint8_t *a, uint8_t *out) { fp_t x, r; memcpy(&x, a, 48); fp_inv(&r, &x); memcpy(out, &r, 48); }

/* ================================================================= Fr / NTT ==================
 * "Next" row f1 of SURVEY.md 8f: ark_poly::Radix2EvaluationDomain<Fr>::{fft,ifft}_in_place and
 * the coset variants with offset Fr::GENERATOR = 7, as legogroth16's witness map uses them
 * (legogroth16/src/r1cs_to_qap.rs:187-207).  ark-poly ^0.4.1 is not vendored; this restates the
 * definition: group generator g = 7^((r-1)/n), evals[i] = sum_j a_j (off * g^i)^j, natural order
 * in and out.  Elements are Montgomery limbs (4 x u64, R = 2^256) as ark-ff stores Fr. */
typedef struct { uint64_t l[4]; } fr_t;
static inline int fr_geq_mod(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) { if (a[i] > FRC_MOD[i]) return 1; if (a[i] < FRC_MOD[i]) return 0; }
    return 1;
}
static inline void fr_sub_mod(uint64_t a[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - FRC_MOD[i] - br; a[i] = (uint64_t)t; br = (t >> 64) & 1; }
}
static inline void fr_add(fr_t *r, const fr_t *a, const fr_t *b) {
    u128 c = 0; uint64_t t[4];
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; t[i] = (uint64_t)c; c >>= 64; }
    if (c || fr_geq_mod(t)) fr_sub_mod(t);
    memcpy(r->l, t, 32);
}
static inline void fr_sub(fr_t *r, const fr_t *a, const fr_t *b) {
    u128 br = 0; uint64_t t[4];
    for (int i = 0; i < 4; i++) { u128 d = (u128)a->l[i] - b->l[i] - br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)t[i] + FRC_MOD[i]; t[i] = (uint64_t)c; c >>= 64; } }
    memcpy(r->l, t, 32);
}
static inline void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) {
    uint64_t t[6] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_INV64;
        c = (u128)m * FRC_MOD[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * FRC_MOD[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || fr_geq_mod(t)) fr_sub_mod(t);
    memcpy(r->l, t, 32);
}
static void fr_pow_u64(fr_t *r, const fr_t *a, uint64_t e) {
    fr_t acc; memcpy(acc.l, FRC_ONE, 32);
    for (int i = 63; i >= 0; i--) { fr_mul(&acc, &acc, &acc); if ((e >> i) & 1) fr_mul(&acc, &acc, a); }
    *r = acc;
}
static void fr_inv(fr_t *r, const fr_t *a) {                 /* a^(r-2) */
    uint64_t e[4]; memcpy(e, FRC_MOD, 32); e[0] -= 2;
    fr_t acc; memcpy(acc.l, FRC_ONE, 32);
    for (int i = 255; i >= 0; i--) { fr_mul(&acc, &acc, &acc); if ((e[i >> 6] >> (i & 63)) & 1) fr_mul(&acc, &acc, a); }
    *r = acc;
}
/* group generator of the size-2^logn domain: W32^(2^(32-logn)) */
static void fr_domain_gen(fr_t *g, uint32_t logn) {
    memcpy(g->l, FRC_W32, 32);
    for (uint32_t i = logn; i < 32; i++) fr_mul(g, g, g);
}
/* in-place radix-2 DIT NTT, natural order in and out, `root` a primitive 2^logn-th root */
static void fr_ntt_core(fr_t *a, uint32_t logn, const fr_t *root) {
    size_t n = (size_t)1 << logn;
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (uint32_t b = 0; b < logn; b++) j |= ((i >> b) & 1) << (logn - 1 - b);
        if (j > i) { fr_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    fr_t *tw = (fr_t *)malloc(sizeof(fr_t) * (n / 2 ? n / 2 : 1));
    memcpy(tw[0].l, FRC_ONE, 32);
    for (size_t i = 1; i < n / 2; i++) fr_mul(&tw[i], &tw[i - 1], root);
    for (uint32_t s = 1; s <= logn; s++) {
        size_t m = (size_t)1 << (s - 1), step = n >> s;
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n / 2; k++) {
            size_t blk = k / m, j = k % m, lo = blk * 2 * m + j, hi = lo + m;
            fr_t v; fr_mul(&v, &a[hi], &tw[j * step]);
            fr_t u = a[lo];
            fr_add(&a[lo], &u, &v); fr_sub(&a[hi], &u, &v);
        }
    }
    free(tw);
}
EXPORT void ref_fr_ntt(uint8_t *data, uint32_t logn, int inverse, int coset) {
    size_t n = (size_t)1 << logn;
    fr_t *a = (fr_t *)data;
    fr_t g, gen; fr_domain_gen(&g, logn); memcpy(gen.l, FRC_GEN, 32);
    if (!inverse) {
        if (coset) { fr_t p; memcpy(p.l, FRC_ONE, 32); for (size_t i = 0; i < n; i++) { fr_mul(&a[i], &a[i], &p); fr_mul(&p, &p, &gen); } }
        fr_ntt_core(a, logn, &g);
    } else {
        fr_t gi; fr_inv(&gi, &g);
        fr_ntt_core(a, logn, &gi);
        fr_t nn, ninv, r2; memset(&nn, 0, sizeof nn); nn.l[0] = n; memcpy(r2.l, FRC_R2, 32);
        fr_mul(&nn, &nn, &r2);                                   /* n in Montgomery form */
        fr_inv(&ninv, &nn);
        fr_t geninv; fr_inv(&geninv, &gen);
        fr_t p = ninv;
        for (size_t i = 0; i < n; i++) { fr_mul(&a[i], &a[i], &p); if (coset) fr_mul(&p, &p, &geninv); }
    }
}
/* tail of LibsnarkReduction::witness_map_from_matrices (legogroth16/src/r1cs_to_qap.rs:187-207) */
EXPORT void ref_qap_h_from_abc(const uint8_t *a_in, const uint8_t *b_in, const uint8_t *c_in, uint32_t logn, uint8_t *out_h) {
    size_t n = (size_t)1 << logn;
    fr_t *a = (fr_t *)malloc(32 * n), *b = (fr_t *)malloc(32 * n), *c = (fr_t *)malloc(32 * n);
    memcpy(a, a_in, 32 * n); memcpy(b, b_in, 32 * n); memcpy(c, c_in, 32 * n);
    fr_t *arr[3] = {a, b, c};
    for (int k = 0; k < 3; k++) { ref_fr_ntt((uint8_t *)arr[k], logn, 1, 0); ref_fr_ntt((uint8_t *)arr[k], logn, 0, 1); }
    fr_t gen, z, one, zinv; memcpy(gen.l, FRC_GEN, 32); memcpy(one.l, FRC_ONE, 32);
    fr_pow_u64(&z, &gen, (uint64_t)n); fr_sub(&z, &z, &one); fr_inv(&zinv, &z);    /* 1 / (7^n - 1) */
    for (size_t i = 0; i < n; i++) { fr_t t; fr_mul(&t, &a[i], &b[i]); fr_sub(&t, &t, &c[i]); fr_mul(&a[i], &t, &zinv); }
    ref_fr_ntt((uint8_t *)a, logn, 1, 1);
    memcpy(out_h, a, 32 * n);
    free(a); free(b); free(c);
}
/* evaluate_constraint over a whole matrix (legogroth16/src/r1cs_to_qap.rs:13-44, :163-186): CSR times assignment,
 * Montgomery Fr in and out; rayon-parallel over the rows in the reference, OpenMP here. */
EXPORT void ref_fr_spmv(const uint32_t *row_ptr, const uint32_t *col, const uint8_t *coeff, size_t rows, const uint8_t *w, uint8_t *out) {
#pragma omp parallel for schedule(static, 1024)
    for (size_t i = 0; i < rows; i++) {
        fr_t acc; memset(&acc, 0, sizeof acc);
        for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
            fr_t c, x, t; memcpy(&c, coeff + 32 * (size_t)k, 32); memcpy(&x, w + 32 * (size_t)col[k], 32);
            fr_mul(&t, &c, &x); fr_add(&acc, &acc, &t);
        }
        memcpy(out + 32 * i, &acc, 32);
    }
}
EXPORT void ref_fr_mul(const uint8_t *a, const uint8_t *b, uint8_t *out) { fr_t x, y, r; memcpy(&x, a, 32); memcpy(&y, b, 32); fr_mul(&r, &x, &y); memcpy(out, &r, 32); }

/* ---- ark-serialize wire format of G1 ("next" row f4; CPU timing baseline and second checker) ----
 * CanonicalDeserialize::deserialize_compressed for a Vec<G1Affine> with Validate::Yes / ::No
 * (utils/src/serde_utils.rs:25-33; legogroth16/src/data_structures.rs:150-168): big-endian x with
 * the Zcash flag bits, y = (x^3 + 4)^((p+1)/4) with the root chosen by the sort flag, subgroup test
 * (beta x, y) == -[x^2] P as ark-bls12-381 0.4 does it (eprint 2021/1130).
 * status: 0 ok, 1 malformed, 2 not on the curve, 3 not in the subgroup. */
EXPORT void ref_g1_deserialize_compressed(const uint8_t *in, size_t n, int validate, uint8_t *out_aff, uint8_t *status) {
    static const uint64_t X2[4] = {0x0000000100000000ULL, 0xac45a4010001a402ULL, 0, 0};     /* x^2 */
    fp_t r2, one_raw, beta, b4;
    memcpy(r2.l, FPC_R2, 48); memcpy(beta.l, FPC_BETA, 48); memcpy(b4.l, FPC_B_G1, 48);
    fp_set_zero(&one_raw); one_raw.l[0] = 1;
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < n; i++) {
        const uint8_t *src = in + 48 * i;
        uint8_t *dst = out_aff + 96 * i;
        memset(dst, 0, 96);
        int comp = src[0] & 0x80, inf = src[0] & 0x40, large = (src[0] & 0x20) != 0;
        if (!comp) { status[i] = 1; continue; }
        fp_t x;
        for (int k = 0; k < 6; k++) {
            uint64_t w = 0;
            for (int j = 0; j < 8; j++) w = (w << 8) | src[8 * (5 - k) + j];
            x.l[k] = w;
        }
        x.l[5] &= 0x1fffffffffffffffULL;
        if (inf) { status[i] = (fp_is_zero(&x) && !large) ? 0 : 1; continue; }
        if (fp_geq_p(x.l)) { status[i] = 1; continue; }
        g1_aff p; p.inf = 0;
        fp_mul(&p.x, &x, &r2);
        fp_t rhs, y, chk, yc;
        fp_sqr(&rhs, &p.x); fp_mul(&rhs, &rhs, &p.x); fp_add(&rhs, &rhs, &b4);
        fp_pow(&y, &rhs, FPC_EXP_SQRT, 6);
        fp_sqr(&chk, &y);
        if (!fp_eq(&chk, &rhs)) { status[i] = 2; continue; }
        fp_mul(&yc, &y, &one_raw);                                   /* canonical y */
        int is_large = 0;
        for (int k = 5; k >= 0; k--) { if (yc.l[k] != FPC_HALF_P[k]) { is_large = yc.l[k] > FPC_HALF_P[k]; break; } }
        if (is_large != large) fp_neg(&y, &y);
        p.y = y;
        if (validate) {
            g1_jac q; g1_mul_bigint(&q, &p, X2);
            g1_aff qa; g1_jac_to_aff(&qa, &q);
            fp_t bx, ny; fp_mul(&bx, &p.x, &beta); fp_neg(&ny, &p.y);
            if (qa.inf || !fp_eq(&qa.x, &bx) || !fp_eq(&qa.y, &ny)) { status[i] = 3; continue; }
        }
        status[i] = 0;
        g1_store(dst, &p);
    }
}

/* G2 counterpart: 96-byte records (x.c1 || x.c0 big-endian), y by the norm-method square root in Fp2,
 * sort flag = lexicographic order on (c1, c0), subgroup test psi(Q) == [x] Q (eprint 2021/1130). */
static int fp_sqrt_c(fp_t *r, const fp_t *a) {
    fp_pow(r, a, FPC_EXP_SQRT, 6);
    fp_t chk; fp_sqr(&chk, r);
    return fp_eq(&chk, a);
}
static int fp_lex_largest_c(const fp_t *mont) {
    fp_t one_raw, c; fp_set_zero(&one_raw); one_raw.l[0] = 1;
    fp_mul(&c, mont, &one_raw);
    for (int k = 5; k >= 0; k--) if (c.l[k] != FPC_HALF_P[k]) return c.l[k] > FPC_HALF_P[k];
    return 0;
}
static int fp2_sqrt_c(fp2_t *out, const fp2_t *a) {
    fp_t s, t;
    if (fp_is_zero(&a->c1)) {
        if (fp_sqrt_c(&s, &a->c0)) { out->c0 = s; fp_set_zero(&out->c1); return 1; }
        fp_neg(&t, &a->c0);
        if (fp_sqrt_c(&s, &t)) { fp_set_zero(&out->c0); out->c1 = s; return 1; }
        return 0;
    }
    fp_t n, n2, two_inv;
    memcpy(two_inv.l, FPC_TWO_INV, 48);
    fp_sqr(&n2, &a->c0); fp_sqr(&t, &a->c1); fp_add(&n2, &n2, &t);
    if (!fp_sqrt_c(&n, &n2)) return 0;
    for (int k = 0; k < 2; k++) {
        fp_t d;
        if (k == 0) fp_add(&d, &a->c0, &n); else fp_sub(&d, &a->c0, &n);
        fp_mul(&d, &d, &two_inv);
        if (!fp_sqrt_c(&s, &d) || fp_is_zero(&s)) continue;
        fp_t s2, inv; fp_add(&s2, &s, &s); fp_inv(&inv, &s2);
        out->c0 = s; fp_mul(&out->c1, &a->c1, &inv);
        fp2_t chk; fp2_sqr(&chk, out);
        if (fp2_eq(&chk, a)) return 1;
    }
    return 0;
}
EXPORT void ref_g2_deserialize_compressed(const uint8_t *in, size_t n, int validate, uint8_t *out_aff, uint8_t *status) {
    static const uint64_t XABS[4] = {0xd201000000010000ULL, 0, 0, 0};
    fp_t r2, b4;
    memcpy(r2.l, FPC_R2, 48); memcpy(b4.l, FPC_B_G1, 48);
    fp2_t psix, psiy, b2;
    memcpy(psix.c0.l, FPC_PSI_X0, 48); memcpy(psix.c1.l, FPC_PSI_X1, 48);
    memcpy(psiy.c0.l, FPC_PSI_Y0, 48); memcpy(psiy.c1.l, FPC_PSI_Y1, 48);
    b2.c0 = b4; b2.c1 = b4;
#pragma omp parallel for schedule(dynamic, 32)
    for (size_t i = 0; i < n; i++) {
        const uint8_t *src = in + 96 * i;
        uint8_t *dst = out_aff + 192 * i;
        memset(dst, 0, 192);
        int comp = src[0] & 0x80, inf = src[0] & 0x40, large = (src[0] & 0x20) != 0;
        if (!comp) { status[i] = 1; continue; }
        fp_t c[2];                                            /* c[0] = x.c1 (first on the wire), c[1] = x.c0 */
        for (int h = 0; h < 2; h++)
            for (int k = 0; k < 6; k++) {
                uint64_t w = 0;
                for (int j = 0; j < 8; j++) w = (w << 8) | src[48 * h + 8 * (5 - k) + j];
                c[h].l[k] = w;
            }
        c[0].l[5] &= 0x1fffffffffffffffULL;
        if (inf) { status[i] = (fp_is_zero(&c[0]) && fp_is_zero(&c[1]) && !large) ? 0 : 1; continue; }
        if (fp_geq_p(c[0].l) || fp_geq_p(c[1].l)) { status[i] = 1; continue; }
        g2_aff p; p.inf = 0;
        fp_mul(&p.x.c1, &c[0], &r2); fp_mul(&p.x.c0, &c[1], &r2);
        fp2_t rhs, y;
        fp2_sqr(&rhs, &p.x); fp2_mul(&rhs, &rhs, &p.x); fp2_add(&rhs, &rhs, &b2);
        if (!fp2_sqrt_c(&y, &rhs)) { status[i] = 2; continue; }
        int is_large = fp_is_zero(&y.c1) ? fp_lex_largest_c(&y.c0) : fp_lex_largest_c(&y.c1);
        if (is_large != large) fp2_neg(&y, &y);
        p.y = y;
        if (validate) {
            g2_jac q; g2_mul_bigint(&q, &p, XABS);
            g2_aff qa; g2_jac_to_aff(&qa, &q);
            fp2_t px, py, t;
            fp2_conj(&t, &p.x); fp2_mul(&px, &t, &psix);
            fp2_conj(&t, &p.y); fp2_mul(&py, &t, &psiy);
            fp2_neg(&py, &py);                                  /* psi(Q) == -[|x|] Q  <=>  [|x|] Q == (psi_x, -psi_y) */
            if (qa.inf || !fp2_eq(&qa.x, &px) || !fp2_eq(&qa.y, &py)) { status[i] = 3; continue; }
        }
        status[i] = 0;
        g2_store(dst, &p);
    }
}
