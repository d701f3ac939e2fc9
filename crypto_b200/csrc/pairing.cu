// Optimal-ate pairing on BLS12-381 for sm_100a: Miller loop + final exponentiation.
//
// Replaces ark_ec::models::bls12::Bls12::{multi_miller_loop, final_exponentiation} as the
// reference reaches them (bbs_plus/src/proof.rs:494, bbs_plus/src/signature.rs:284,
// legogroth16/src/verifier.rs:69-80, vb_accumulator/src/positive.rs:420,
// utils/src/randomized_pairing_check.rs:134,169,204-214).
//
// Design: a pairing is a long dependent chain of Fp12 operations, so one thread per pairing
// would leave the chip idle and take tens of milliseconds.  Instead one 128-thread CTA is a
// "tower engine": operands live in shared memory, and every Fp12 multiplication is spread over
// the CTA - 108 lanes each do ONE Fp Montgomery multiplication (36 Fp2 products x 3 Karatsuba
// parts over the flattened basis Fp12 = Fp2[w]/(w^6 - xi)), 12 lanes recombine.  The G2
// doubling/addition steps are issued as two waves of independent Fp2 products the same way.
// The grid has one CTA per pair; partial Miller values are multiplied by a tree of CTAs and
// one CTA runs the final exponentiation (eprint 2020/875 chain, the arkworks convention:
// the result is e(P,Q)^3 w.r.t. the textbook exponent (p^12-1)/r).
//
// Line functions follow ark-ec 0.4 bls12::G2Prepared (homogeneous projective doubling /
// addition, M-type twist, ell = mul_by_014(c0, c1*P.x, c2*P.y)), so the Miller-loop value itself
// is bit-identical to arkworks', not only the pairing.
#include "common.cuh"
#include "ec.cuh"
#include "fp_inv.cuh"

namespace dg {

#define PAIR_THREADS 128
#define BLS_X_ABS 0xd201000000010000ULL

// Logical worker id of a thread: workers 0, 1, 2, 3 are lane 0 of warps 0, 1, 2, 3, workers 4 .. 7 lane 1, and so on.
// The engine's small linear stages give different workers different jobs ("worker 0: e = ..., worker 1: h = ...");
// as lanes of ONE warp those jobs would run one after the other (divergence), as lanes of different warps they run
// side by side on the four schedulers.  Every function below indexes its work by this id only.
__device__ __forceinline__ int pair_wid() { return (int)((threadIdx.x & 31u) * 4u + ((threadIdx.x >> 5) & 3u)); }
// An engine is a GROUP of 128 threads (4 warps); a CTA holds one group (most kernels) or two (k_miller: one group walks
// the G2 point and produces the lines, the other consumes them into f).  Every engine function synchronises its own
// group only, on named barrier 1 + group.
__device__ __forceinline__ void pair_sync() { asm volatile("bar.sync %0, 128;" ::"r"(1u + (threadIdx.x >> 7)) : "memory"); }

// Fp12 in shared/global memory: 12 Fp in ark order (c0.c0.c0, c0.c0.c1, c0.c1.c0, ..., c1.c2.c1).
// Flattened basis: coefficient of w^k (an Fp2) lives at tower position (i = k&1, j = k>>1).
struct F12 { Fp c[12]; };
__device__ __forceinline__ int widx(int k) { return (k & 1) * 6 + (k >> 1) * 2; }

// operands by value: see fp_mul_ni in fp2.cuh (with pointer parameters the a * b rows are not fused into IMAD.WIDE)
static __device__ __noinline__ Fp fp_mul_val(Fp a, Fp b) { return fp_mul(a, b); }
static __device__ __forceinline__ Fp fp_mul_smem(const Fp *a, const Fp *b) { return fp_mul_val(*a, *b); }

// One Karatsuba part of an Fp2 product: 0: a0*b0, 1: a1*b1, 2: (a0+a1)*(b0+b1)
// ONE call site for the multiplier: with three calls in three branches the lanes of a warp holding parts 0, 1 and 2
// would run the (out-of-line, ~600-instruction) multiplication three times in a row.
__device__ __forceinline__ void fp2_part_ops(int part, const Fp *a, const Fp *b, Fp &x, Fp &y) {
    Fp a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    Fp sa = fp_add(a0, a1), sb = fp_add(b0, b1);
    x = fsel(part == 2, sa, fsel(part == 1, a1, a0));
    y = fsel(part == 2, sb, fsel(part == 1, b1, b0));
}
__device__ __forceinline__ Fp fp2_part(int part, const Fp *a, const Fp *b) {
    Fp x, y;
    fp2_part_ops(part, a, b, x, y);
    return fp_mul_val(x, y);
}
// recombine three parts into the Fp2 product
__device__ __forceinline__ void fp2_from_parts(Fp *d, const Fp *p) {
    Fp c0 = fp_sub(p[0], p[1]);
    Fp c1 = fp_sub(fp_sub(p[2], p[0]), p[1]);
    d[0] = c0; d[1] = c1;
}

struct Engine {
    Fp prod[108];      // product scratch (3 Karatsuba parts of up to 36 Fp2 products)
    Fp q[72];          // the 36 Fp2 products after recombination / xi-twist
    Fp tmp[24];        // linear-stage scratch
};

// a / 2 mod p without a multiplication: make it even by adding p, then shift right
__device__ __forceinline__ Fp fp_half(const Fp &a) {
    uint32_t odd = 0u - (a.l[0] & 1u);
    Fp t;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(t.l[0]) : "r"(a.l[0]), "r"(odd & fp_p_limb(0)));
#pragma unroll
    for (int i = 1; i < 12; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(t.l[i]) : "r"(a.l[i]), "r"(odd & fp_p_limb(i)));
    Fp r;
#pragma unroll
    for (int i = 0; i < 11; i++) r.l[i] = (t.l[i] >> 1) | (t.l[i + 1] << 31);
    r.l[11] = t.l[11] >> 1;                 // a + p < 2^382: no carry out of limb 11
    return r;
}

// Unreduced accumulation for the linear stages: integers below 8p fit the 12 limbs (8p < 2^384), so up to six reduced
// values are added as plain integers (one 12-limb carry chain each, no trial subtraction) and brought back below p by
// three conditional subtractions of 4p, 2p, p -- about half the instructions of six modular additions, all of them on
// the stage's critical path.
// (fp_add_raw: fp.cuh)
template <int K> __device__ __forceinline__ void fp_csub_kp(Fp &a) {          // a -= K p when a >= K p  (K = 1, 2, 4)
    uint32_t t[12], br;
    constexpr int SH = K == 4 ? 2 : K == 2 ? 1 : 0;
    auto kp = [](int i) -> uint32_t {
        uint64_t lo = i ? fp_p_limb(i - 1) : 0u, hi = fp_p_limb(i);
        return SH ? (uint32_t)(((hi << 32) | lo) >> (32 - SH)) : (uint32_t)hi;
    };
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(a.l[0]), "r"(kp(0)));
#pragma unroll
    for (int i = 1; i < 12; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[i]) : "r"(a.l[i]), "r"(kp(i)));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(br));
    const bool keep = br != 0;
#pragma unroll
    for (int i = 0; i < 12; i++) a.l[i] = keep ? a.l[i] : t[i];
}
__device__ __forceinline__ Fp fp_reduce_lt8p(Fp a) {
    fp_csub_kp<4>(a);
    fp_csub_kp<2>(a);
    fp_csub_kp<1>(a);
    return a;
}

// C = A * B (C may alias A or B).  All PAIR_THREADS threads must call.
//   phase 1: 108 lanes, one Fp multiplication each (36 Fp2 products x 3 Karatsuba parts)
//   phase 2:  72 lanes, one per component of the 36 Fp2 products, recombine the Karatsuba parts and apply the w^6 = xi
//             twist when i + j >= 6 (two modular subtractions each)
//   phase 3:  12 lanes (degree k, component) add the six products of their column unreduced and reduce once
__device__ void f12_mul(Engine &e, F12 *C, const F12 *A, const F12 *B) {
    int tid = pair_wid();
    if (tid < 108) {
        int pr = tid / 3, part = tid - pr * 3, i = pr / 6, j = pr - i * 6;
        e.prod[tid] = fp2_part(part, &A->c[widx(i)], &B->c[widx(j)]);
    }
    pair_sync();
    if (tid < 72) {                                       // one lane per COMPONENT of the 36 products (comp is warp-uniform)
        int pr = tid >> 1, comp = tid & 1, i = pr / 6, j = pr - i * 6;
        const Fp *p = &e.prod[pr * 3];
        const bool tw = i + j >= 6;                       // times xi = 1 + u:  (c0 - c1, c0 + c1) = (2 p0 - p2, p2 - 2 p1)
        Fp out;
        if (comp == 0) out = fp_sub(fsel(tw, fp_dbl(p[0]), p[0]), fsel(tw, p[2], p[1]));
        else out = fp_sub(fp_sub(p[2], fsel(tw, p[1], p[0])), p[1]);
        e.q[tid] = out;
    }
    pair_sync();
    if (tid < 12) {
        int k = tid >> 1, comp = tid & 1;
        Fp acc = e.q[2 * (0 * 6 + k) + comp];             // i = 0, j = k
#pragma unroll
        for (int i = 1; i < 6; i++) {
            int j = k - i; if (j < 0) j += 6;
            acc = fp_add_raw(acc, e.q[2 * (i * 6 + j) + comp]);
        }
        C->c[widx(k) + comp] = fp_reduce_lt8p(acc);        // six reduced terms: < 6p
    }
    pair_sync();
}

// C = A * L for a line value L = l0 + l1 w^2 + l2 w^3 (ark's mul_by_014 on the tower = positions w^0, w^2, w^3 of the
// flattened basis), L given as three Fp2 (six Fp).  Same three phases as f12_mul over the 18 non-zero products; the
// results are canonical field elements, hence identical to f12_mul with the line embedded in a full Fp12.
__device__ void f12_mul_line(Engine &e, F12 *C, const F12 *A, const Fp *L) {
    int tid = pair_wid();
    if (tid < 54) {
        int pr = tid / 3, part = tid - pr * 3, i = pr / 3, jj = pr - i * 3;
        e.prod[tid] = fp2_part(part, &A->c[widx(i)], &L[2 * jj]);
    }
    pair_sync();
    if (tid < 36) {
        int pr = tid >> 1, comp = tid & 1, i = pr / 3, jj = pr - i * 3, j = jj ? jj + 1 : 0;
        const Fp *p = &e.prod[pr * 3];
        const bool tw = i + j >= 6;
        Fp out;
        if (comp == 0) out = fp_sub(fsel(tw, fp_dbl(p[0]), p[0]), fsel(tw, p[2], p[1]));
        else out = fp_sub(fp_sub(p[2], fsel(tw, p[1], p[0])), p[1]);
        e.q[tid] = out;
    }
    pair_sync();
    if (tid < 12) {
        int k = tid >> 1, comp = tid & 1;
        int i0 = k, i1 = k - 2 < 0 ? k + 4 : k - 2, i2 = k - 3 < 0 ? k + 3 : k - 3;            // i + j = k (mod 6) for j = 0, 2, 3
        Fp acc = fp_add_raw(fp_add_raw(e.q[2 * (i0 * 3 + 0) + comp], e.q[2 * (i1 * 3 + 1) + comp]), e.q[2 * (i2 * 3 + 2) + comp]);
        C->c[widx(k) + comp] = fp_reduce_lt8p(acc);
    }
    pair_sync();
}

// C = A^2 for A in the cyclotomic subgroup (every value after the easy part of the final
// exponentiation, and every GT element): Granger-Scott squaring.  With Fp4 = Fp2[t]/(t^2 - xi),
// t = w^3, write A = X + Y w + Z w^2, X = (a0, a3), Y = (a1, a4), Z = (a2, a5); then
//   A^2 = (3 X^2 - 2 conj X) + (3 t Z^2 + 2 conj Y) w + (3 Y^2 - 2 conj Z) w^2
// (checked against the generic square in the big-integer oracle).  Nine Fp2 squarings = 18 Fp
// multiplications in ONE wave of 18 lanes, then 12 lanes recombine.  C may alias A.
__device__ void f12_cyc_sqr(Engine &e, F12 *C, const F12 *A) {
    int tid = pair_wid();
    if (tid < 18) {
        // lane = 2 s + h;  s = 3 m + j:  Fp4 number m (0: X, 1: Y, 2: Z), j = 0: x0, 1: x1, 2: x0 + x1
        int sidx = tid >> 1, h = tid & 1, m = sidx / 3, j = sidx - 3 * m;
        const Fp *x0 = &A->c[widx(m)], *x1 = &A->c[widx(m + 3)];
        Fp v0, v1;
        if (j == 0) { v0 = x0[0]; v1 = x0[1]; }
        else if (j == 1) { v0 = x1[0]; v1 = x1[1]; }
        else { v0 = fp_add(x0[0], x1[0]); v1 = fp_add(x0[1], x1[1]); }
        // complex squaring (v0 + v1 u)^2 = (v0 + v1)(v0 - v1) + 2 v0 v1 u : lane h takes one product
        Fp a = h ? v0 : fp_add(v0, v1), b = h ? v1 : fp_sub(v0, v1);
        e.prod[tid] = fp_mul_smem(&a, &b);
    }
    pair_sync();
    if (tid < 12) {
        int k = tid >> 1, comp = tid & 1;
        // component `c` of the Fp2 square number s:  c = 0: prod[2s],  c = 1: 2 * prod[2s + 1]
        auto S = [&](int sq, int c) -> Fp { return c ? fp_dbl(e.prod[2 * sq + 1]) : e.prod[2 * sq]; };
        // Fp4 square of number m:  part 0 = S0 + xi S1,  part 1 = S2 - S0 - S1
        auto part0 = [&](int m, int c) -> Fp {
            Fp s1a = S(3 * m + 1, 0), s1b = S(3 * m + 1, 1);
            Fp x = c ? fp_add(s1a, s1b) : fp_sub(s1a, s1b);               // (xi S1).c
            return fp_add(S(3 * m, c), x);
        };
        auto part1 = [&](int m, int c) -> Fp { return fp_sub(fp_sub(S(3 * m + 2, c), S(3 * m, c)), S(3 * m + 1, c)); };
        // Even k = 2m takes part 0 of the Fp4 square number m (minus sign); odd k takes part 1 of number
        // (k == 1 ? 2 : (k - 3) / 2) (plus sign), times xi for k == 1.  With the worker numbering above the even and the
        // odd outputs sit in different warps, and inside each branch the lanes differ only in data (m, comp).
        Fp v;
        const bool plus = (k & 1) != 0;
        if (!plus) {
            v = part0(k >> 1, comp);
        } else {
            const int m = k == 1 ? 2 : (k - 3) >> 1;
            Fp p0 = part1(m, 0), p1 = part1(m, 1);
            Fp tw = comp ? fp_add(p0, p1) : fp_sub(p0, p1);                 // (xi * part1).comp
            v = fsel(k == 1, tw, comp ? p1 : p0);
        }
        Fp three = fp_add(fp_dbl(v), v), two_a = fp_dbl(A->c[widx(k) + comp]);
        C->c[widx(k) + comp] = plus ? fp_add(three, two_a) : fp_sub(three, two_a);
    }
    pair_sync();
}

__device__ void f12_copy(F12 *d, const F12 *s) {
    int tid = pair_wid();
    if (tid < 12) d->c[tid] = s->c[tid];
    pair_sync();
}
__device__ void f12_set_one(F12 *d) {
    int tid = pair_wid();
    if (tid < 12) d->c[tid] = tid == 0 ? fp_one() : fp_zero();
    pair_sync();
}
// conjugation over Fp6 (= p^6 Frobenius): negate the w-odd half (tower c1 = indices 6..11)
__device__ void f12_conj(F12 *d, const F12 *s) {
    int tid = pair_wid();
    if (tid < 12) d->c[tid] = tid < 6 ? s->c[tid] : fp_neg(s->c[tid]);
    pair_sync();
}
// d = s^(p^pw), pw in {1,2,3}: coefficient of w^k -> conj^pw(a_k) * xi^(k (p^pw - 1)/6)
__device__ void f12_frobenius(Engine &e, F12 *d, const F12 *s, int pw) {
    int tid = pair_wid();
    if (tid < 18) {
        int k = tid / 3, part = tid - k * 3;
        Fp a[2] = {s->c[widx(k)], s->c[widx(k) + 1]};
        if (pw & 1) a[1] = fp_neg(a[1]);
        const uint32_t(*g)[2][12] = pw == 1 ? DGC_FROB1 : pw == 2 ? DGC_FROB2 : DGC_FROB3;
        Fp b[2];
#pragma unroll
        for (int t = 0; t < 12; t++) { b[0].l[t] = g[k][0][t]; b[1].l[t] = g[k][1][t]; }
        e.prod[tid] = fp2_part(part, a, b);
    }
    pair_sync();
    if (tid < 6) fp2_from_parts(&d->c[widx(tid)], &e.prod[tid * 3]);
    pair_sync();
}

// Fp inversion by one thread (a^(p-2)); the only long serial chain in the final exponentiation
static __device__ __noinline__ Fp fp_inv_serial(Fp a) {
    Fp tbl[16];
    tbl[0] = fp_one();
    tbl[1] = a;
    for (int i = 2; i < 16; i++) tbl[i] = fp_mul_smem(&tbl[i - 1], &a);
    uint32_t ex[12];
#pragma unroll
    for (int i = 0; i < 12; i++) ex[i] = fp_p_limb(i);
    ex[0] -= 2;
    Fp acc = fp_one();
    for (int nib = 95; nib >= 0; nib--) {
        for (int k = 0; k < 4; k++) acc = fp_mul_smem(&acc, &acc);
        uint32_t d = (ex[nib >> 3] >> ((nib & 7) * 4)) & 15;
        acc = fp_mul_smem(&acc, &tbl[d]);
    }
    return acc;
}

// d = 1/s via norms:  abar = conj(s);  N = s*abar in Fp6;  N^-1 = sigma(N) sigma^2(N) / Norm_{Fp6/Fp2}(N)
// with sigma = p^2 Frobenius;  the Fp2 norm is inverted with one Fp inversion.
__device__ void f12_inv(Engine &e, F12 *d, const F12 *s, F12 *t0, F12 *t1, F12 *t2) {
    int tid = pair_wid();
    f12_conj(t0, s);                 // abar
    f12_mul(e, t1, s, t0);           // N (odd half is zero)
    f12_frobenius(e, t2, t1, 2);     // sigma(N)
    f12_mul(e, t0, t0, t2);          // abar * sigma(N)
    f12_frobenius(e, t2, t2, 2);     // sigma^2(N)
    f12_mul(e, t0, t0, t2);          // abar * sigma(N) * sigma^2(N)
    f12_frobenius(e, t2, t2, 2);     // back to N  (sigma^3 = id on Fp6)
    f12_frobenius(e, d, t2, 2);      // sigma(N)
    f12_mul(e, t1, t2, d);           // N sigma(N)
    f12_frobenius(e, d, d, 2);       // sigma^2(N)
    f12_mul(e, t1, t1, d);           // Norm in Fp2: only the w^0 coefficient is non-zero
    if (tid == 0) {
        Fp x = t1->c[0], y = t1->c[1];
        Fp n = fp_inv_pornin(fp_add(fp_mul_val(x, x), fp_mul_val(y, y)));
        t1->c[0] = fp_mul_val(x, n);
        t1->c[1] = fp_neg(fp_mul_val(y, n));
    }
    pair_sync();
    f12_mul(e, d, t0, t1);
}

// d = s^|x| conjugated (x < 0): ark Bls12::exp_by_x.  d must not alias s.
__device__ void f12_exp_by_x(Engine &e, F12 *d, const F12 *s) {
    f12_copy(d, s);                                   // top bit (63) of |x|
    for (int i = 62; i >= 0; i--) {
        f12_cyc_sqr(e, d, d);                          // operands of exp_by_x are cyclotomic
        if ((BLS_X_ABS >> i) & 1) f12_mul(e, d, d, s);
    }
    f12_conj(d, d);
}

// Bls12::final_exponentiation, same operation order as arkworks; r = in/out, 5 temporaries.
__device__ void f12_final_exp(Engine &e, F12 *r, F12 *f1, F12 *f2, F12 *y0, F12 *y1, F12 *y2) {
    f12_conj(f1, r);
    f12_inv(e, f2, r, y0, y1, y2);
    f12_mul(e, r, f1, f2);                 // f^(p^6-1)
    f12_copy(f2, r);
    f12_frobenius(e, r, r, 2);
    f12_mul(e, r, r, f2);                  // ^(p^2+1)
    f12_cyc_sqr(e, y0, r);                 // y0 = r^2 (ark: cyclotomic_square)
    f12_exp_by_x(e, y1, r);
    f12_conj(y2, r);
    f12_mul(e, y1, y1, y2);
    f12_exp_by_x(e, y2, y1);
    f12_conj(y1, y1);
    f12_mul(e, y1, y1, y2);
    f12_exp_by_x(e, y2, y1);
    f12_frobenius(e, y1, y1, 1);
    f12_mul(e, y1, y1, y2);
    f12_mul(e, r, r, y0);
    f12_exp_by_x(e, y0, y1);
    f12_exp_by_x(e, y2, y0);
    f12_frobenius(e, y0, y1, 2);
    f12_conj(y1, y1);
    f12_mul(e, y1, y1, y2);
    f12_mul(e, y1, y1, y0);
    f12_mul(e, r, r, y1);
}

// ---- Miller loop --------------------------------------------------------------------------------
struct MillerState {
    Fp rx[2], ry[2], rz[2];     // running G2 point, homogeneous projective
    Fp qx[2], qy[2];            // Q affine
    Fp px, py;                  // P affine
    Fp co[3][2];                // line coefficients of the current step
};

__device__ __forceinline__ void fp2s_add(Fp *d, const Fp *a, const Fp *b) { Fp x = fp_add(a[0], b[0]), y = fp_add(a[1], b[1]); d[0] = x; d[1] = y; }
__device__ __forceinline__ void fp2s_sub(Fp *d, const Fp *a, const Fp *b) { Fp x = fp_sub(a[0], b[0]), y = fp_sub(a[1], b[1]); d[0] = x; d[1] = y; }
__device__ __forceinline__ void fp2s_neg(Fp *d, const Fp *a) { Fp x = fp_neg(a[0]), y = fp_neg(a[1]); d[0] = x; d[1] = y; }
__device__ __forceinline__ void fp2s_half(Fp *d, const Fp *a) { Fp x = fp_half(a[0]), y = fp_half(a[1]); d[0] = x; d[1] = y; }

// Doubling step (ark G2Prepared double_in_place): two product waves; the linear algebra between
// them is spread over several lanes instead of one.
//   T layout (Fp2 = 2 slots): 0 m1=rx*ry  2 b  4 c  6 j  8 s  10 e  12 f  14 a  16 g  18 h  20 b-f
__device__ void miller_double(Engine &e, MillerState &m, Fp *line) {
    int tid = pair_wid();
    Fp *T = e.tmp;
    // wave 1: 0: rx*ry  1: ry^2  2: rz^2  3: rx^2  4: (ry+rz)^2
    if (tid < 15) {
        int job = tid / 3, part = tid - job * 3;
        Fp sum[2];
        const Fp *A, *B;
        switch (job) {
            case 0: A = m.rx; B = m.ry; break;
            case 1: A = m.ry; B = m.ry; break;
            case 2: A = m.rz; B = m.rz; break;
            case 3: A = m.rx; B = m.rx; break;
            default: fp2s_add(sum, m.ry, m.rz); A = sum; B = sum; break;
        }
        e.prod[tid] = fp2_part(part, A, B);
    }
    pair_sync();
    // The linear stages work per Fp2 COMPONENT (additions, halvings and negations are component-wise; xi mixes the two
    // inputs but each output component is still one lane's job), and the workers are placed so that the long job sits
    // alone in its warp: workers 0/1 (warps 0/1) own the e, f chain, workers 2/3, 6/7, 10/11 (warps 2/3) the rest.
    if (tid < 10) {                                                        // m1, b, c, j, s from their Karatsuba parts
        int w = tid >> 1, comp = tid & 1;
        const Fp *p = &e.prod[3 * w];
        T[2 * w + comp] = comp ? fp_sub(fp_sub(p[2], p[0]), p[1]) : fp_sub(p[0], p[1]);
    }
    pair_sync();
    if (tid < 2) {                                                         // e = (4+4u)*3c = 4*xi*3c ; f = 3e
        Fp d = tid ? fp_add(T[4], T[5]) : fp_sub(T[4], T[5]);              // (xi c).comp
        Fp d3 = fp_add(fp_dbl(d), d);
        Fp ee = fp_dbl(fp_dbl(d3));
        T[10 + tid] = ee;
        T[12 + tid] = fp_add(fp_dbl(ee), ee);
    } else if (tid == 2 || tid == 3) {                                     // h = s - (b + c) ; co2 = -h
        int comp = tid - 2;
        Fp h = fp_sub(T[8 + comp], fp_add(T[2 + comp], T[4 + comp]));
        T[18 + comp] = h;
        m.co[2][comp] = fp_neg(h);
    } else if (tid == 6 || tid == 7) {                                     // co1 = 3j
        int comp = tid - 6;
        Fp j = T[6 + comp];
        m.co[1][comp] = fp_add(fp_dbl(j), j);
    } else if (tid == 10 || tid == 11) {                                   // a = m1 / 2
        int comp = tid - 10;
        T[14 + comp] = fp_half(T[comp]);
    }
    pair_sync();
    if (tid < 2) {                                                         // g = (b + f) / 2
        T[16 + tid] = fp_half(fp_add(T[2 + tid], T[12 + tid]));
    } else if (tid == 2 || tid == 3) {                                     // co0 = i = e - b
        int comp = tid - 2;
        m.co[0][comp] = fp_sub(T[10 + comp], T[2 + comp]);
    } else if (tid == 6 || tid == 7) {                                     // b - f
        int comp = tid - 6;
        T[20 + comp] = fp_sub(T[2 + comp], T[12 + comp]);
    }
    pair_sync();
    // wave 2: 0: a*(b-f)  1: g^2  2: e^2  3: b*h, and on four more workers the line value of this step:
    // line = (co0, co1 * px, co2 * py)  (ell(coeffs, P) of ark's Miller loop).  One call site for all 16 multiplications.
    if (tid < 16) {
        Fp x, y;
        if (tid < 12) {
            int job = tid / 3, part = tid - job * 3;
            const Fp *A, *B;
            switch (job) {
                case 0: A = &T[14]; B = &T[20]; break;
                case 1: A = &T[16]; B = &T[16]; break;
                case 2: A = &T[10]; B = &T[10]; break;
                default: A = &T[2]; B = &T[18]; break;
            }
            fp2_part_ops(part, A, B, x, y);
        } else {
            const int j = tid & 1, hi = tid >= 14;
            x = hi ? m.co[2][j] : m.co[1][j];
            y = hi ? m.py : m.px;
        }
        Fp r = fp_mul_val(x, y);
        if (tid < 12) e.prod[tid] = r; else line[tid - 10] = r;             // line[2..3] = co1 px, line[4..5] = co2 py
    } else if (tid < 18) {
        line[tid - 16] = m.co[0][tid - 16];
    }
    pair_sync();
    if (tid < 2) {                                                         // ry = g^2 - 3 e^2, one component per worker
        const Fp *pg = &e.prod[3], *pe = &e.prod[6];
        Fp g2 = tid ? fp_sub(fp_sub(pg[2], pg[0]), pg[1]) : fp_sub(pg[0], pg[1]);
        Fp e2 = tid ? fp_sub(fp_sub(pe[2], pe[0]), pe[1]) : fp_sub(pe[0], pe[1]);
        m.ry[tid] = fp_sub(g2, fp_add(fp_dbl(e2), e2));
    } else if (tid == 2 || tid == 3) {                                     // rx = a (b - f)
        int comp = tid - 2;
        const Fp *p = &e.prod[0];
        m.rx[comp] = comp ? fp_sub(fp_sub(p[2], p[0]), p[1]) : fp_sub(p[0], p[1]);
    } else if (tid == 6 || tid == 7) {                                     // rz = b h
        int comp = tid - 6;
        const Fp *p = &e.prod[9];
        m.rz[comp] = comp ? fp_sub(fp_sub(p[2], p[0]), p[1]) : fp_sub(p[0], p[1]);
    }
    pair_sync();
}

// Addition step (ark G2Prepared add_in_place), three product waves.
__device__ void miller_add(Engine &e, MillerState &m, Fp *line) {
    int tid = pair_wid();
    Fp *T = e.tmp;   // 0: theta 2: lambda 4: c 6: d 8: e 10: f 12: g 14: h / (g-h)
    // wave 1: qy*rz, qx*rz
    if (tid < 6) {
        int job = tid / 3, part = tid - job * 3;
        e.prod[tid] = fp2_part(part, job == 0 ? m.qy : m.qx, m.rz);
    }
    pair_sync();
    if (tid == 0) {
        Fp t[2];
        fp2_from_parts(t, &e.prod[0]); fp2s_sub(&T[0], m.ry, t);      // theta
        fp2_from_parts(t, &e.prod[3]); fp2s_sub(&T[2], m.rx, t);      // lambda
    }
    pair_sync();
    // wave 2: c = theta^2, d = lambda^2, theta*qx, lambda*qy
    if (tid < 12) {
        int job = tid / 3, part = tid - job * 3;
        const Fp *A, *B;
        switch (job) {
            case 0: A = &T[0]; B = &T[0]; break;
            case 1: A = &T[2]; B = &T[2]; break;
            case 2: A = &T[0]; B = m.qx; break;
            default: A = &T[2]; B = m.qy; break;
        }
        e.prod[tid] = fp2_part(part, A, B);
    }
    pair_sync();
    if (tid == 0) {
        Fp a[2], b[2], t[2];
        fp2_from_parts(&T[4], &e.prod[0]);        // c
        fp2_from_parts(&T[6], &e.prod[3]);        // d
        fp2_from_parts(a, &e.prod[6]); fp2_from_parts(b, &e.prod[9]);
        fp2s_sub(t, a, b);                        // j
        m.co[0][0] = t[0]; m.co[0][1] = t[1];
        fp2s_neg(t, &T[0]);
        m.co[1][0] = t[0]; m.co[1][1] = t[1];
        m.co[2][0] = T[2]; m.co[2][1] = T[3];
    }
    pair_sync();
    // wave 3: e = lambda*d, f = rz*c, g = rx*d, and the line value of this step on four more workers
    if (tid < 13) {
        Fp x, y;
        if (tid < 9) {
            int job = tid / 3, part = tid - job * 3;
            const Fp *A, *B;
            switch (job) {
                case 0: A = &T[2]; B = &T[6]; break;
                case 1: A = m.rz; B = &T[4]; break;
                default: A = m.rx; B = &T[6]; break;
            }
            fp2_part_ops(part, A, B, x, y);
        } else {
            const int j = (tid - 9) & 1, hi = tid >= 11;
            x = hi ? m.co[2][j] : m.co[1][j];
            y = hi ? m.py : m.px;
        }
        Fp r = fp_mul_val(x, y);
        if (tid < 9) e.prod[tid] = r; else line[tid - 7] = r;               // line[2..3] = co1 px, line[4..5] = co2 py
    } else if (tid < 15) {
        line[tid - 13] = m.co[0][tid - 13];
    }
    pair_sync();
    if (tid == 0) {
        Fp t[2];
        fp2_from_parts(&T[8], &e.prod[0]); fp2_from_parts(&T[10], &e.prod[3]); fp2_from_parts(&T[12], &e.prod[6]);
        fp2s_add(t, &T[8], &T[10]); fp2s_sub(t, t, &T[12]); fp2s_sub(&T[14], t, &T[12]);   // h = e + f - 2g
    }
    pair_sync();
    // wave 4: lambda*h, theta*(g-h), e*ry, rz*e
    if (tid == 0) { Fp t[2]; fp2s_sub(t, &T[12], &T[14]); T[12] = t[0]; T[13] = t[1]; }
    pair_sync();
    if (tid < 12) {
        int job = tid / 3, part = tid - job * 3;
        const Fp *A, *B;
        switch (job) {
            case 0: A = &T[2]; B = &T[14]; break;
            case 1: A = &T[0]; B = &T[12]; break;
            case 2: A = &T[8]; B = m.ry; break;
            default: A = m.rz; B = &T[8]; break;
        }
        e.prod[tid] = fp2_part(part, A, B);
    }
    pair_sync();
    if (tid == 0) {
        Fp a[2], b[2];
        fp2_from_parts(m.rx, &e.prod[0]);
        fp2_from_parts(a, &e.prod[3]); fp2_from_parts(b, &e.prod[6]);
        fp2s_sub(m.ry, a, b);
        fp2_from_parts(m.rz, &e.prod[9]);
    }
    pair_sync();
}

struct PairSmem {
    Engine e;
    MillerState m;
    F12 f, t[5];
};

// Miller loop, one CTA of TWO engine groups per pair.  The walk of the G2 point (doubling / addition steps and their line
// coefficients) does not depend on f, so group 1 runs it on its own and leaves the 68 line values
// (co0, co1 * px, co2 * py) in a shared-memory ring; group 0 only squares f and multiplies the lines in
// (f <- f^2 * line: two Fp12 operations per step instead of those two plus the point step's two or three product waves
// and the line scaling).  The producer never waits; the consumer spins on a counter the producer bumps after each line.
#define MILLER_LINES 68                          // 63 doubling steps + 5 addition steps (bits of |x| below the top one)
struct MillerSmem {
    Engine ec, ep;                               // consumer / producer scratch
    MillerState m;
    F12 f;
    Fp ring[MILLER_LINES * 6];
    int produced;
    int skip;
};
__device__ __forceinline__ void miller_publish(MillerSmem &S, int n) {        // after the step's last pair_sync
    if (pair_wid() == 0) {
        __threadfence_block();
        *(volatile int *)&S.produced = n;
    }
}
__device__ __forceinline__ void miller_wait(MillerSmem &S, int s) {
    while (*(volatile int *)&S.produced <= s) {}
    __threadfence_block();
}
// out[pair] = Miller value (before the final conjugation), 1 for identity pairs.
__global__ void __launch_bounds__(2 * PAIR_THREADS) k_miller(const Affine<Fp> *g1, const Affine<Fp2> *g2, uint32_t k, F12 *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MillerSmem &S = *reinterpret_cast<MillerSmem *>(smem_raw);
    const int tid = pair_wid(), group = threadIdx.x >> 7;
    uint32_t pair = blockIdx.x;
    if (threadIdx.x == 0) {
        Affine<Fp> p = aff_load<Fp>(&g1[pair]);
        Affine<Fp2> q = aff_load<Fp2>(&g2[pair]);
        S.skip = aff_is_inf(p) || aff_is_inf(q);
        S.produced = 0;
        S.m.px = p.x; S.m.py = p.y;
        S.m.qx[0] = q.x.c0; S.m.qx[1] = q.x.c1; S.m.qy[0] = q.y.c0; S.m.qy[1] = q.y.c1;
        S.m.rx[0] = q.x.c0; S.m.rx[1] = q.x.c1; S.m.ry[0] = q.y.c0; S.m.ry[1] = q.y.c1;
        S.m.rz[0] = fp_one(); S.m.rz[1] = fp_zero();
    }
    if (group == 0) f12_set_one(&S.f);
    __syncthreads();                                                       // both groups: state and skip flag are in place
    if (!S.skip) {
        int s = 0;
        if (group == 1) {
            for (int i = 62; i >= 0; i--) {
                miller_double(S.ep, S.m, &S.ring[6 * s]);
                miller_publish(S, ++s);
                if ((BLS_X_ABS >> i) & 1) {
                    miller_add(S.ep, S.m, &S.ring[6 * s]);
                    miller_publish(S, ++s);
                }
            }
        } else {
            for (int i = 62; i >= 0; i--) {
                f12_mul(S.ec, &S.f, &S.f, &S.f);
                miller_wait(S, s);
                f12_mul_line(S.ec, &S.f, &S.f, &S.ring[6 * s]);
                s++;
                if ((BLS_X_ABS >> i) & 1) {
                    miller_wait(S, s);
                    f12_mul_line(S.ec, &S.f, &S.f, &S.ring[6 * s]);
                    s++;
                }
            }
        }
    }
    if (group == 0 && tid < 12) fp_store(&out[pair].c[tid], S.f.c[tid]);
}

// out[b] = product of in[b*8 .. b*8+8)
__global__ void __launch_bounds__(PAIR_THREADS) k_f12_reduce8(const F12 *in, uint32_t n, F12 *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairSmem &S = *reinterpret_cast<PairSmem *>(smem_raw);
    int tid = pair_wid();
    uint32_t lo = blockIdx.x * 8, hi = lo + 8 < n ? lo + 8 : n;
    if (tid < 12) S.f.c[tid] = fp_load_rw(&in[lo].c[tid]);
    pair_sync();
    for (uint32_t i = lo + 1; i < hi; i++) {
        if (tid < 12) S.t[0].c[tid] = fp_load_rw(&in[i].c[tid]);
        pair_sync();
        f12_mul(S.e, &S.f, &S.f, &S.t[0]);
    }
    if (tid < 12) fp_store(&out[blockIdx.x].c[tid], S.f.c[tid]);
}

// mode bit 0: conjugate first (Miller loop epilogue, x < 0); bit 1: final exponentiation;
// in == nullptr means "one".  flags[0] = 1 iff the input to the final exponentiation was zero,
// flags[1] = 1 iff the result equals one.
__global__ void __launch_bounds__(PAIR_THREADS) k_f12_finish(const F12 *in, int mode, F12 *out, int32_t *flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairSmem &S = *reinterpret_cast<PairSmem *>(smem_raw);
    int tid = pair_wid();
    __shared__ int is_zero;
    if (in) { if (tid < 12) S.f.c[tid] = fp_load_rw(&in->c[tid]); pair_sync(); }
    else f12_set_one(&S.f);
    if (mode & 1) f12_conj(&S.f, &S.f);
    if (tid == 0) {
        int z = 1;
        for (int i = 0; i < 12; i++) z &= fp_is_zero(S.f.c[i]);
        is_zero = z;
    }
    pair_sync();
    if ((mode & 2) && !is_zero) f12_final_exp(S.e, &S.f, &S.t[0], &S.t[1], &S.t[2], &S.t[3], &S.t[4]);
    if (tid < 12) fp_store(&out->c[tid], S.f.c[tid]);
    if (tid == 0 && flags) {
        flags[0] = (mode & 2) ? is_zero : 0;
        int one = fp_eq(S.f.c[0], fp_one());
        for (int i = 1; i < 12; i++) one &= fp_is_zero(S.f.c[i]);
        flags[1] = one;
    }
}

// out = a * b   /   out = a^scalar (255-bit canonical integer, MSB first square-and-multiply)
__global__ void __launch_bounds__(PAIR_THREADS) k_f12_mul_or_pow(const F12 *a, const F12 *b, const uint32_t *scalar, F12 *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairSmem &S = *reinterpret_cast<PairSmem *>(smem_raw);
    int tid = pair_wid();
    if (tid < 12) S.t[0].c[tid] = fp_load_rw(&a->c[tid]);
    pair_sync();
    if (b) {
        if (tid < 12) S.t[1].c[tid] = fp_load_rw(&b->c[tid]);
        pair_sync();
        f12_mul(S.e, &S.f, &S.t[0], &S.t[1]);
    } else {
        f12_set_one(&S.f);
        // GT elements are cyclotomic, so the squarings use Granger-Scott; the accumulator starts at 1
        for (int i = 255; i >= 0; i--) {
            f12_cyc_sqr(S.e, &S.f, &S.f);
            if ((scalar[i >> 5] >> (i & 31)) & 1) f12_mul(S.e, &S.f, &S.f, &S.t[0]);
        }
    }
    if (tid < 12) fp_store(&out->c[tid], S.f.c[tid]);
}

static int32_t pairing_smem_opt_in() {
    int32_t rc = smem_opt_in(k_miller, sizeof(MillerSmem));
    if (!rc) rc = smem_opt_in(k_f12_reduce8, sizeof(PairSmem));
    if (!rc) rc = smem_opt_in(k_f12_finish, sizeof(PairSmem));
    if (!rc) rc = smem_opt_in(k_f12_mul_or_pow, sizeof(PairSmem));
    return rc;
}

// mode: 1 = Miller loop only (conjugated), 3 = Miller + final exponentiation
static int32_t pairing_run(const uint8_t *g1, const uint8_t *g2, size_t k, int mode, uint8_t *out_fp12, int32_t *flags_out) {
    int32_t rc = check_init();
    if (rc) return rc;
    if ((k && (!g1 || !g2)) || (!out_fp12 && !flags_out)) return fail(DG_ERR_BAD_ARG, "pairing: null pointer");
    rc = pairing_smem_opt_in();
    if (rc) return rc;
    ThreadState &t = tls();
    size_t need = Arena::pad(96 * k) + Arena::pad(192 * k) + 2 * Arena::pad(sizeof(F12) * (k + 1)) + Arena::pad(sizeof(F12)) + 256;
    rc = t.arena.ensure(need, t.stream);
    if (rc) return rc;
    Affine<Fp> *d_p = t.arena.alloc<Affine<Fp>>(k ? k : 1);
    Affine<Fp2> *d_q = t.arena.alloc<Affine<Fp2>>(k ? k : 1);
    F12 *buf0 = t.arena.alloc<F12>(k + 1), *buf1 = t.arena.alloc<F12>(k + 1), *d_out = t.arena.alloc<F12>(1);
    int32_t *d_flags = t.arena.alloc<int32_t>(2);
    const F12 *cur = nullptr;
    if (k) {
        DG_CUDA(cudaMemcpyAsync(d_p, g1, 96 * k, cudaMemcpyHostToDevice, t.stream));
        DG_CUDA(cudaMemcpyAsync(d_q, g2, 192 * k, cudaMemcpyHostToDevice, t.stream));
        DG_LAUNCH(k_miller, (unsigned)k, 2 * PAIR_THREADS, sizeof(MillerSmem), t.stream, d_p, d_q, (uint32_t)k, buf0);
        size_t n = k;
        F12 *src = buf0, *dst = buf1;
        while (n > 1) {
            unsigned nb = div_up(n, 8);
            DG_LAUNCH(k_f12_reduce8, nb, PAIR_THREADS, sizeof(PairSmem), t.stream, src, (uint32_t)n, dst);
            F12 *tmp = src; src = dst; dst = tmp;
            n = nb;
        }
        cur = src;
    }
    DG_LAUNCH(k_f12_finish, 1, PAIR_THREADS, sizeof(PairSmem), t.stream, cur, mode, d_out, d_flags);
    int32_t flags[2] = {0, 0};
    if (out_fp12) DG_CUDA(cudaMemcpyAsync(out_fp12, d_out, sizeof(F12), cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaMemcpyAsync(t.err_flag_host + 8, d_flags, 8, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    flags[0] = ((int32_t *)(t.err_flag_host + 8))[0];
    flags[1] = ((int32_t *)(t.err_flag_host + 8))[1];
    if (flags_out) { flags_out[0] = flags[0]; flags_out[1] = flags[1]; }
    return DG_OK;
}

// Several independent pairing products at once (SnarkPack's GIPA rounds issue six per round,
// legogroth16/src/aggregation/utils.rs:85-97): one k_miller launch covers every pair of every
// product, then each product's CTA tree + final exponentiation runs on its own stream so the
// latency-bound tails overlap instead of queueing behind each other.
#define DG_PAIR_STREAMS 8
static int32_t pairing_batch_run(const uint8_t *g1, const uint8_t *g2, const size_t *counts, size_t nbatch, uint8_t *out_fp12) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!counts || !out_fp12 || nbatch == 0) return fail(DG_ERR_BAD_ARG, "multi_pairing_batch: null pointer");
    size_t k = 0;
    for (size_t i = 0; i < nbatch; i++) k += counts[i];
    if (k && (!g1 || !g2)) return fail(DG_ERR_BAD_ARG, "multi_pairing_batch: null pointer");
    rc = pairing_smem_opt_in();
    if (rc) return rc;
    ThreadState &t = tls();
    static thread_local cudaStream_t aux[DG_PAIR_STREAMS] = {nullptr};
    static thread_local cudaEvent_t ev_fork = nullptr, ev_join[DG_PAIR_STREAMS] = {nullptr};
    if (!ev_fork) {
        DG_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < DG_PAIR_STREAMS; i++) {
            DG_CUDA(cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking));
            DG_CUDA(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
        }
    }
    size_t need = Arena::pad(96 * k) + Arena::pad(192 * k) + 2 * Arena::pad(sizeof(F12) * (k + nbatch)) + Arena::pad(sizeof(F12) * nbatch) + 256;
    rc = t.arena.ensure(need, t.stream);
    if (rc) return rc;
    Affine<Fp> *d_p = t.arena.alloc<Affine<Fp>>(k ? k : 1);
    Affine<Fp2> *d_q = t.arena.alloc<Affine<Fp2>>(k ? k : 1);
    F12 *buf0 = t.arena.alloc<F12>(k + nbatch), *buf1 = t.arena.alloc<F12>(k + nbatch), *d_out = t.arena.alloc<F12>(nbatch);
    if (k) {
        DG_CUDA(cudaMemcpyAsync(d_p, g1, 96 * k, cudaMemcpyHostToDevice, t.stream));
        DG_CUDA(cudaMemcpyAsync(d_q, g2, 192 * k, cudaMemcpyHostToDevice, t.stream));
        DG_LAUNCH(k_miller, (unsigned)k, 2 * PAIR_THREADS, sizeof(MillerSmem), t.stream, d_p, d_q, (uint32_t)k, buf0);
    }
    DG_CUDA(cudaEventRecord(ev_fork, t.stream));
    size_t off = 0;
    for (size_t b = 0; b < nbatch; b++) {
        cudaStream_t s = aux[b % DG_PAIR_STREAMS];
        DG_CUDA(cudaStreamWaitEvent(s, ev_fork, 0));
        size_t n = counts[b];
        F12 *src = buf0 + off, *dst = buf1 + off;
        const F12 *cur = n ? src : nullptr;
        while (n > 1) {
            unsigned nb = div_up(n, 8);
            DG_LAUNCH(k_f12_reduce8, nb, PAIR_THREADS, sizeof(PairSmem), s, src, (uint32_t)n, dst);
            F12 *tmp = src; src = dst; dst = tmp;
            n = nb;
            cur = src;
        }
        DG_LAUNCH(k_f12_finish, 1, PAIR_THREADS, sizeof(PairSmem), s, cur, 3, d_out + b, (int32_t *)nullptr);
        off += counts[b];
    }
    for (size_t b = 0; b < nbatch && b < DG_PAIR_STREAMS; b++) {
        DG_CUDA(cudaEventRecord(ev_join[b], aux[b]));
        DG_CUDA(cudaStreamWaitEvent(t.stream, ev_join[b], 0));
    }
    DG_CUDA(cudaMemcpyAsync(out_fp12, d_out, sizeof(F12) * nbatch, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

}  // namespace dg

using namespace dg;

extern "C" {

int32_t dg_multi_pairing_batch(const uint8_t *g1, const uint8_t *g2, const size_t *counts, size_t nbatch, uint8_t *out_fp12) {
    return pairing_batch_run(g1, g2, counts, nbatch, out_fp12);
}

int32_t dg_multi_miller_loop(const uint8_t *g1, const uint8_t *g2, size_t k, uint8_t *out_fp12) {
    if (!out_fp12) return fail(DG_ERR_BAD_ARG, "multi_miller_loop: null output");
    return pairing_run(g1, g2, k, 1, out_fp12, nullptr);
}
int32_t dg_multi_pairing(const uint8_t *g1, const uint8_t *g2, size_t k, uint8_t *out_fp12) {
    if (!out_fp12) return fail(DG_ERR_BAD_ARG, "multi_pairing: null output");
    return pairing_run(g1, g2, k, 3, out_fp12, nullptr);
}
int32_t dg_multi_pairing_is_one(const uint8_t *g1, const uint8_t *g2, size_t k, int32_t *result) {
    if (!result) return fail(DG_ERR_BAD_ARG, "multi_pairing_is_one: null output");
    int32_t flags[2];
    int32_t rc = pairing_run(g1, g2, k, 3, nullptr, flags);
    if (rc) return rc;
    *result = flags[1];
    return DG_OK;
}

int32_t dg_final_exponentiation(const uint8_t *in_fp12, uint8_t *out_fp12, int32_t *is_some) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!in_fp12 || !out_fp12 || !is_some) return fail(DG_ERR_BAD_ARG, "final_exponentiation: null pointer");
    rc = pairing_smem_opt_in();
    if (rc) return rc;
    ThreadState &t = tls();
    rc = t.arena.ensure(2 * Arena::pad(sizeof(F12)) + 256, t.stream);
    if (rc) return rc;
    F12 *d_in = t.arena.alloc<F12>(1), *d_out = t.arena.alloc<F12>(1);
    int32_t *d_flags = t.arena.alloc<int32_t>(2);
    DG_CUDA(cudaMemcpyAsync(d_in, in_fp12, sizeof(F12), cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_f12_finish, 1, PAIR_THREADS, sizeof(PairSmem), t.stream, d_in, 2, d_out, d_flags);
    DG_CUDA(cudaMemcpyAsync(out_fp12, d_out, sizeof(F12), cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaMemcpyAsync(t.err_flag_host + 8, d_flags, 8, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    *is_some = ((int32_t *)(t.err_flag_host + 8))[0] ? 0 : 1;
    return DG_OK;
}

static int32_t f12_binary(const uint8_t *a, const uint8_t *b, const uint8_t *scalar, uint8_t *out) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!a || !out || (!b && !scalar)) return fail(DG_ERR_BAD_ARG, "fp12 op: null pointer");
    rc = pairing_smem_opt_in();
    if (rc) return rc;
    ThreadState &t = tls();
    rc = t.arena.ensure(3 * Arena::pad(sizeof(F12)) + 256, t.stream);
    if (rc) return rc;
    F12 *d_a = t.arena.alloc<F12>(1), *d_b = t.arena.alloc<F12>(1), *d_out = t.arena.alloc<F12>(1);
    uint32_t *d_s = t.arena.alloc<uint32_t>(8);
    DG_CUDA(cudaMemcpyAsync(d_a, a, sizeof(F12), cudaMemcpyHostToDevice, t.stream));
    if (b) DG_CUDA(cudaMemcpyAsync(d_b, b, sizeof(F12), cudaMemcpyHostToDevice, t.stream));
    else DG_CUDA(cudaMemcpyAsync(d_s, scalar, 32, cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_f12_mul_or_pow, 1, PAIR_THREADS, sizeof(PairSmem), t.stream, d_a, b ? d_b : (const F12 *)nullptr, d_s, d_out);
    DG_CUDA(cudaMemcpyAsync(out, d_out, sizeof(F12), cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}
int32_t dg_gt_pow(const uint8_t *in_fp12, const uint8_t *scalar, uint8_t *out_fp12) { return f12_binary(in_fp12, nullptr, scalar, out_fp12); }
int32_t dg_fp12_mul(const uint8_t *a, const uint8_t *b, uint8_t *out) { return f12_binary(a, b, nullptr, out); }

}  // extern "C"
