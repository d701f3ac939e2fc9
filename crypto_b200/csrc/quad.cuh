// Quad-cooperative XYZZ arithmetic for the latency-bound phases of the MSM (bucket reduction,
// window combination).  Those phases are long chains of dependent point additions / doublings
// with little parallelism across points, so a thread-per-point mapping leaves the SM idle while
// each thread grinds through 14 dependent Montgomery multiplications (~1850 cycles each, measured
// with tools/fpmul_bench.cu).  Here 4 adjacent lanes own ONE point operation: the 14 (add) or 9
// (double) field multiplications are issued as 4 (3) waves of up to four independent products,
// one product per lane, with operands and results exchanged through a per-quad shared-memory
// workspace.  All lanes run the same instruction stream (operands are selected by pointer), so a
// warp of 8 quads never diverges on the common path.
#pragma once
#include "ec.cuh"

namespace dg {

// workspace slots: object k (an XYZZ point) lives in v[4k .. 4k+3] = X, Y, ZZ, ZZZ
#define DG_Q_OBJS 4          // run, acc, item, spare
#define DG_Q_TMP (4 * DG_Q_OBJS)
template <class F> struct QuadWS { F v[DG_Q_TMP + 16]; };

struct QuadCtx {
    uint32_t ql;        // lane within the quad, 0..3
    uint32_t mask;      // __syncwarp mask of this quad
};
__device__ __forceinline__ QuadCtx quad_ctx() {
    uint32_t lane = threadIdx.x & 31;
    QuadCtx q;
    q.ql = lane & 3;
    q.mask = 0xFu << (lane & ~3u);
    return q;
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
    fstore(reinterpret_cast<char *>(dst) + q.ql * sizeof(F), w.v[4 * S + q.ql]);
}

// D = 2 * A (dbl-2008-s-1, a = 0); A must not be the identity.  D may alias A.
// Operands are selected per lane BY VALUE (a short divergent copy), then one convergent fmul.
template <class F> __device__ __noinline__ void quad_dbl(QuadWS<F> &w, int D, int A, QuadCtx q) {
    F *T = &w.v[DG_Q_TMP];
    const F *PA = &w.v[4 * A];                             // X, Y, ZZ, ZZZ
    F a, b, r;
    // wave 1: V = (2Y)^2 (even lanes), XX = X^2 (odd lanes)
    if (q.ql & 1) a = PA[0]; else a = fdbl(PA[1]);
    r = fmul(a, a);
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
    r = fmul(a, b);
    T[3 + q.ql] = r;                                       // T3 = W, T4 = S, T5 = MM, T6 = ZZ3
    __syncwarp(q.mask);
    F X3 = fsub(T[5], fdbl(T[4]));
    // wave 3: M*(S - X3), W*Y, W*ZZZ
    switch (q.ql) {
        case 0: a = M; b = fsub(T[4], X3); break;
        case 1: a = T[3]; b = PA[1]; break;
        default: a = T[3]; b = PA[3]; break;
    }
    r = fmul(a, b);
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
    T[q.ql] = fmul(a, b);
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
    T[4 + q.ql] = fmul(a, b);
    __syncwarp(q.mask);
    // wave 3: PPP = P*PP, Q = U1*PP, ZZ3 = ZZ12*PP
    switch (q.ql) {
        case 0: a = P; break;
        case 1: a = T[0]; break;
        default: a = T[6]; break;
    }
    r = fmul(a, T[4]);
    if (q.ql < 3) T[8 + q.ql] = r;                         // T8 = PPP, T9 = Q, T10 = ZZ3
    __syncwarp(q.mask);
    F X3 = fsub(fsub(T[5], T[8]), fdbl(T[9]));
    // wave 4: R*(Q - X3), S1*PPP, ZZZ12*PPP
    switch (q.ql) {
        case 0: a = R; b = fsub(T[9], X3); break;
        case 1: a = T[2]; b = T[8]; break;
        default: a = T[7]; b = T[8]; break;
    }
    r = fmul(a, b);
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
