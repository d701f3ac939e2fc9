// Quad-cooperative XYZZ arithmetic for the latency-bound phases of the MSM (bucket reduction,
// window combination).  Those phases are long chains of dependent point additions / doublings
// with little parallelism across points, so a thread-per-point mapping leaves the SM idle while
// each thread grinds through 14 dependent Montgomery multiplications (~1850 cycles each, measured
// with tools/fpmul_bench.cu).  Here 4 adjacent lanes own ONE point operation: the 14 (add) or 9
// (double) field multiplications are issued as 4 (3) waves of up to four independent products,
// one product per lane, with operands and results exchanged through a per-quad shared-memory
// workspace.  All lanes run the same instruction stream (operands are selected by pointer), so a
// warp of 8 quads never diverges on the common path.
//
// Over Fp2 (G2) one product is three Fp multiplications (Karatsuba), which one lane would run back to back, so there a
// "quad" is 12 lanes: each of the four quad members is a trio of adjacent lanes holding the same operands, trio lane k
// multiplies Karatsuba part k (a0 b0, a1 b1, (a0 + a1)(b0 + b1)) and the three products are exchanged by shuffles.
// A wave then costs one Fp multiplication plus ~36 shuffles instead of three multiplications and the call overhead
// (measured on B200, G2 MSM at 2^18 terms: window combination 1.89 -> see DESIGN.md section 3).  Two such quads fit a warp;
// lanes 24..31 idle.
#pragma once
#include "ec.cuh"

namespace dg {

// workspace slots: object k (an XYZZ point) lives in v[4k .. 4k+3] = X, Y, ZZ, ZZZ; the 14 temporaries of an addition follow.
// Two objects (accumulator, item) + temporaries = 22 field elements per quad: 33 KB per CTA of 32 (G1) / 16 (G2) quads, so
// the register file (5 CTAs/SM at 96 registers), not shared memory, bounds the residency of the line-sum kernel.
#define DG_Q_OBJS 2          // acc, item
#define DG_Q_ACC 0
#define DG_Q_ITEM 1
#define DG_Q_TMP (4 * DG_Q_OBJS)
template <class F> struct QuadWS { F v[DG_Q_TMP + 14]; };

struct QuadCtx {
    uint32_t ql;        // member within the quad, 0..3
    uint32_t mask;      // __syncwarp / __shfl_sync mask of this quad
    uint32_t part;      // Fp2: Karatsuba part of this lane within its trio (0..2); Fp: 0
    uint32_t base;      // Fp2: first lane of this lane's trio
    uint32_t qi;        // index of the quad within the CTA
    bool active;        // false for the warp's left-over lanes (Fp2: lanes 24..31): they only take part in CTA barriers
    bool wide;          // 12-lane quads (Fp2 only)
};
// WIDE = 12 lanes per quad.  It pays where a launch is latency-bound (few CTAs, long dependent chains); a launch that
// fills the GPU with CTAs is better off with 4-lane quads (no idle lanes, no redundant linear work in the trios).
template <class F> struct QuadWide { static constexpr bool value = sizeof(F) > 48; };
template <bool WIDE> struct QuadLanes {
    static constexpr int LPQ = WIDE ? 12 : 4;                 // lanes per quad
    static constexpr int QPW = 32 / LPQ;                      // quads per warp
    static constexpr int cta_threads(int quads) { return (quads + QPW - 1) / QPW * 32; }
};
template <bool WIDE> __device__ __forceinline__ QuadCtx quad_ctx() {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QuadCtx q;
    q.wide = WIDE;
    if (WIDE) {
        const uint32_t qw = lane / 12, l12 = lane - 12 * qw;
        q.ql = l12 / 3;
        q.part = l12 - 3 * q.ql;
        q.base = 12 * qw + 3 * q.ql;
        q.active = qw < 2;
        q.mask = q.active ? (0xFFFu << (12 * qw)) : 0xFF000000u;
        q.qi = warp * 2 + (q.active ? qw : 0);
    } else {
        q.ql = lane & 3;
        q.part = 0;
        q.base = lane;
        q.active = true;
        q.mask = 0xFu << (lane & ~3u);
        q.qi = threadIdx.x >> 2;
    }
    return q;
}

// one product per quad member: plain for Fp, split over the member's trio for Fp2 (all 12 lanes of the quad call this
// together; the operands are identical in the three lanes of a trio)
__device__ __forceinline__ Fp qmul(const Fp &a, const Fp &b, const QuadCtx &) { return fmul(a, b); }
__device__ __forceinline__ Fp2 qmul(const Fp2 &a, const Fp2 &b, const QuadCtx &q) {
    if (!q.wide) return fmul(a, b);
    Fp x = fsel(q.part == 0, a.c0, a.c1), y = fsel(q.part == 0, b.c0, b.c1);
    Fp xs = fp_add(a.c0, a.c1), ys = fp_add(b.c0, b.c1);
    x = fsel(q.part == 2, xs, x);
    y = fsel(q.part == 2, ys, y);
    Fp t = fp_mul_ni(x, y), t0, t1, t2;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        t0.l[i] = __shfl_sync(q.mask, t.l[i], q.base);
        t1.l[i] = __shfl_sync(q.mask, t.l[i], q.base + 1);
        t2.l[i] = __shfl_sync(q.mask, t.l[i], q.base + 2);
    }
    Fp2 r;
    r.c0 = fp_sub(t0, t1);
    r.c1 = fp_sub(fp_sub(t2, t0), t1);
    return r;
}

template <class F> __device__ __forceinline__ void quad_set_inf(QuadWS<F> &w, int D, const QuadCtx &q) {
    w.v[4 * D + q.ql] = fzero<F>();
    __syncwarp(q.mask);
}
template <class F> __device__ __forceinline__ void quad_load(QuadWS<F> &w, int D, const XYZZ<F> *src, const QuadCtx &q) {
    w.v[4 * D + q.ql] = fload_rw<F>(reinterpret_cast<const char *>(src) + q.ql * sizeof(F));
    __syncwarp(q.mask);
}
template <class F> __device__ __forceinline__ void quad_store(const QuadWS<F> &w, int S, XYZZ<F> *dst, const QuadCtx &q) {
    if (q.part == 0) fstore(reinterpret_cast<char *>(dst) + q.ql * sizeof(F), w.v[4 * S + q.ql]);
}

// D = 2 * A (dbl-2008-s-1, a = 0); A must not be the identity.  D may alias A.
// Operands are selected per lane BY VALUE (a short divergent copy), then one convergent fmul.
template <class F> __device__ __noinline__ void quad_dbl(QuadWS<F> &w, int D, int A, QuadCtx q) {
    F *T = &w.v[DG_Q_TMP];
    const F *PA = &w.v[4 * A];                             // X, Y, ZZ, ZZZ
    F a, b, r;
    // wave 1: V = (2Y)^2 (even lanes), XX = X^2 (odd lanes)
    if (q.ql & 1) a = PA[0]; else a = fdbl(PA[1]);
    r = qmul(a, a, q);
    if (q.ql < 2) T[q.ql] = r;                             // T0 = V, T1 = XX
    if (q.ql == 2) T[2] = a;                               // T2 = U = 2Y
    __syncwarp(q.mask);
    F M = fadd(fdbl(T[1]), T[1]);                          // 3 X^2
    // wave 2: W = U*V, S = X*V, MM = M^2, ZZ3 = V*ZZ
    switch (q.ql) {
        case 0: a = T[2]; b = T[0]; break;
        case 1: a = PA[0]; b = T[0]; break;
        case 2: a = M; b = M; break;
        default: a = T[0]; b = PA[2]; break;
    }
    r = qmul(a, b, q);
    T[3 + q.ql] = r;                                       // T3 = W, T4 = S, T5 = MM, T6 = ZZ3
    __syncwarp(q.mask);
    F X3 = fsub(T[5], fdbl(T[4]));
    // wave 3: M*(S - X3), W*Y, W*ZZZ
    switch (q.ql) {
        case 0: a = M; b = fsub(T[4], X3); break;
        case 1: a = T[3]; b = PA[1]; break;
        default: a = T[3]; b = PA[3]; break;
    }
    r = qmul(a, b, q);
    if (q.ql < 3) T[7 + q.ql] = r;                         // T7, T8, T9 = ZZZ3
    __syncwarp(q.mask);
    F out;
    switch (q.ql) {
        case 0: out = X3; break;
        case 1: out = fsub(T[7], T[8]); break;
        case 2: out = T[6]; break;
        default: out = T[9]; break;
    }
    __syncwarp(q.mask);
    w.v[4 * D + q.ql] = out;
    __syncwarp(q.mask);
}

// D = A + B (add-2008-s), complete: identity operands, A == B, A == -B.  D may alias A or B.
template <class F> __device__ __noinline__ void quad_add(QuadWS<F> &w, int D, int A, int B, QuadCtx q) {
    F *T = &w.v[DG_Q_TMP];
    const F *PA = &w.v[4 * A], *PB = &w.v[4 * B];
    bool a_inf = fis_zero(PA[2]), b_inf = fis_zero(PB[2]);
    if (a_inf || b_inf) {                                  // quad-uniform
        F t = a_inf ? PB[q.ql] : PA[q.ql];
        __syncwarp(q.mask);
        w.v[4 * D + q.ql] = t;
        __syncwarp(q.mask);
        return;
    }
    F a, b, r;
    // wave 1: U1 = X1*ZZ2, U2 = X2*ZZ1, S1 = Y1*ZZZ2, S2 = Y2*ZZZ1
    switch (q.ql) {
        case 0: a = PA[0]; b = PB[2]; break;
        case 1: a = PB[0]; b = PA[2]; break;
        case 2: a = PA[1]; b = PB[3]; break;
        default: a = PB[1]; b = PA[3]; break;
    }
    T[q.ql] = qmul(a, b, q);
    __syncwarp(q.mask);
    F P = fsub(T[1], T[0]), R = fsub(T[3], T[2]);
    if (fis_zero(P)) {                                     // quad-uniform, rare
        __syncwarp(q.mask);                                // every lane has read T[0..3] before quad_dbl reuses them
        if (fis_zero(R)) {
            quad_dbl(w, D, A, q);
        } else {
            __syncwarp(q.mask);
            w.v[4 * D + q.ql] = fzero<F>();
            __syncwarp(q.mask);
        }
        return;
    }
    // wave 2: PP = P^2, RR = R^2, ZZ12 = ZZ1*ZZ2, ZZZ12 = ZZZ1*ZZZ2
    switch (q.ql) {
        case 0: a = P; b = P; break;
        case 1: a = R; b = R; break;
        case 2: a = PA[2]; b = PB[2]; break;
        default: a = PA[3]; b = PB[3]; break;
    }
    T[4 + q.ql] = qmul(a, b, q);
    __syncwarp(q.mask);
    // wave 3: PPP = P*PP, Q = U1*PP, ZZ3 = ZZ12*PP
    switch (q.ql) {
        case 0: a = P; break;
        case 1: a = T[0]; break;
        default: a = T[6]; break;
    }
    r = qmul(a, T[4], q);
    if (q.ql < 3) T[8 + q.ql] = r;                         // T8 = PPP, T9 = Q, T10 = ZZ3
    __syncwarp(q.mask);
    F X3 = fsub(fsub(T[5], T[8]), fdbl(T[9]));
    // wave 4: R*(Q - X3), S1*PPP, ZZZ12*PPP
    switch (q.ql) {
        case 0: a = R; b = fsub(T[9], X3); break;
        case 1: a = T[2]; b = T[8]; break;
        default: a = T[7]; b = T[8]; break;
    }
    r = qmul(a, b, q);
    if (q.ql < 3) T[11 + q.ql] = r;                        // T11, T12, T13 = ZZZ3
    __syncwarp(q.mask);
    F out;
    switch (q.ql) {
        case 0: out = X3; break;
        case 1: out = fsub(T[11], T[12]); break;
        case 2: out = T[10]; break;
        default: out = T[13]; break;
    }
    __syncwarp(q.mask);
    w.v[4 * D + q.ql] = out;
    __syncwarp(q.mask);
}

}  // namespace dg
