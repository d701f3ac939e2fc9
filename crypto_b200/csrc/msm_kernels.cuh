// Pippenger multi-scalar multiplication on sm_100a, templated over the base field (G1: Fp,
// G2: Fp2).  Replaces ark-ec 0.4 VariableBaseMSM::msm_bigint as the reference calls it
// (legogroth16/src/prover.rs:286,299,363,592; bbs_plus/src/setup.rs:145;
// schnorr_pok/src/pok_generalized_pedersen.rs:97,153; vb_accumulator/src/witness.rs:415; ...
// SURVEY.md 8a rows a4/a5).
//
// Pipeline (all on one stream, no host round trip):
//   1 k_digits<COUNT>   signed radix-2^c digits of every scalar, per-(window,bucket) histogram
//   2 k_scan_*          exclusive prefix sum of the histogram -> bucket offsets
//   3 k_digits<SCATTER> counting-sort the (point index, sign) entries by (window, bucket)
//   4 k_accumulate      perfectly balanced segmented accumulation: every thread folds a
//                       fixed-length chunk of the sorted entry list into XYZZ partial sums
//                       (run time independent of the scalar distribution - witnesses full of
//                       0/1 values do not serialise on hot buckets)
//   5 k_bucket_fixup    stitches the partial sums of buckets that straddle chunks; hot buckets go to
//                       k_fixup_long_part / k_fixup_long_final (split over up to 32 CTAs of quads)
//   6 k_red_lines / k_red_subsets / k_red_final
//                       sum_b (b+1)*B[w][b] per window from row / column sums and weighted subset sums
//   7 k_window_combine  Horner over windows, result as Jacobian (ark Projective layout)
#pragma once
#include "ec.cuh"
#include "glv.cuh"
#include "quad.cuh"

namespace dg {

struct MsmGeom {
    int c;            // window bits
    int ndig;         // signed digits per scalar (msm_ndigits): the top digit never overflows
    int nwin;         // bucket sets: ndig, or 1 when the bases carry precomputed 2^(c*k) multiples
    uint32_t nbw;     // buckets per set = 2^(c-1)
    uint32_t nb;      // total buckets = nwin * nbw
    uint32_t row_stride;   // precomputed bases: row k (= 2^(c*k) * P_i) starts at k * row_stride; 0 = plain bases
    int fp2;               // G2 (host-side heuristics only)
    int glv;               // plain bases: every scalar is split s = +-k1 - k2 * lambda (k1, k2 < 2^127), ndig digits each
    uint32_t phi_off;      // GLV: phi(P_i) = (beta x_i, y_i) is stored at index phi_off + i of the expanded base array
    int w0;                // plain bases: this run covers the digit positions [w0, w0 + nwin) only (window-group split,
                           // msm_host.cuh msm_run); 0 with nwin == ndig is the whole scalar
};

// ---------------------------------------------------------------- digits / counting sort -----
template <int PASS>   // 0 = count, 1 = scatter
__global__ void __launch_bounds__(256) k_digits(const uint32_t *__restrict__ scalars, uint32_t n, MsmGeom g,
                                                uint32_t *__restrict__ counters, uint32_t *__restrict__ entries,
                                                uint32_t *__restrict__ err_flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *sp = reinterpret_cast<const uint4 *>(scalars) + 2 * (size_t)i;
    uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    uint32_t s[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0};
    // Halve the range: s > (r - 1) / 2 is replaced by r - s with every digit's sign flipped ([s]P = [r - s](-P)), so
    // the recoded integer is < 2^254 and ceil(254 / c) digits suffice (15 instead of 16 at c = 17).  A scalar >= r
    // borrows in r - s: not canonical, flagged.
    bool flip = false;
    {
        constexpr uint32_t RM[8] = {DG_R0, DG_R1, DG_R2, DG_R3, DG_R4, DG_R5, DG_R6, DG_R7};
        uint32_t t[8], borrow = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {                      // t = r - s
            uint64_t d = (uint64_t)RM[k] - s[k] - borrow;
            t[k] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
        bool zero = true, t_less = false;                  // t < s  <=>  s > r - s
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            if (zero && t[k] != s[k]) { t_less = t[k] < s[k]; zero = false; }
        }
        if (borrow) {
            if (PASS == 0) atomicOr(err_flag, 1u);         // s > r (s == r gives t = 0 and is caught below as s >= r too)
            return;
        }
        bool t_is_zero = true;
#pragma unroll
        for (int k = 0; k < 8; k++) t_is_zero &= t[k] == 0;
        if (t_is_zero) {                                   // s == r
            if (PASS == 0) atomicOr(err_flag, 1u);
            return;
        }
        if (t_less) {
            flip = true;
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = t[k];
        }
    }
    const uint32_t mask = (1u << g.c) - 1, half = 1u << (g.c - 1);
    // signed radix-2^c digits of the value v[0 .. words) for the point at index `pidx`; `flip_sign` negates every digit
    auto emit = [&](const uint32_t *v, int words, uint32_t pidx, bool flip_sign) {
        uint32_t carry = 0;
        const int w_end = g.row_stride ? g.ndig : g.w0 + g.nwin;   // digits below w0 only feed the carry
        for (int w = 0; w < w_end; w++) {
            int bit = w * g.c;
            uint32_t raw = 0;
            if (bit < 32 * words) {
                int word = bit >> 5, off = bit & 31;
                uint64_t two = ((uint64_t)(word + 1 < words ? v[word + 1] : 0u) << 32) | v[word];
                raw = (uint32_t)(two >> off) & mask;
            }
            uint32_t d = raw + carry;
            bool neg = d > half;
            carry = neg ? 1u : 0u;
            uint32_t mag = neg ? (1u << g.c) - d : d;
            if (mag != 0 && w >= g.w0) {
                // precomputed rows fold every digit position into ONE bucket set: digit w of scalar i
                // selects the point 2^(c*w) * P_i stored at w * row_stride + i
                uint32_t bucket = (g.row_stride ? 0u : (uint32_t)(w - g.w0) * g.nbw) + mag - 1;
                if (PASS == 0) {
                    atomicAdd(&counters[bucket], 1u);
                } else {
                    uint32_t pos = atomicAdd(&counters[bucket], 1u);
                    entries[pos] = (pidx + (uint32_t)w * g.row_stride) | ((neg != flip_sign) ? 0x80000000u : 0u);
                }
            }
        }
        if (PASS == 0 && carry && w_end == g.ndig) atomicOr(err_flag, 1u);   // cannot happen for a canonical scalar (msm_ndigits / glv_ndigits)
    };
    if (g.glv) {
        uint32_t k1[4], k2[4];
        bool neg1;
        glv_split(s, k1, neg1, k2);
        emit(k1, 4, i, flip != neg1);                      // [s] P = sigma ( +-[k1] P - [k2] phi(P) ),  sigma = -1 when flipped
        emit(k2, 4, g.phi_off + i, !flip);
    } else {
        emit(s, 8, i, flip);
    }
}

// exclusive scan of `in[0..n)` into `out[0..n]` (out[n] = total); 3 small kernels, 4096 items/block
#define DG_SCAN_ITEMS 4096
// blockIdx.y selects one of several independent arrays (the per-round counts of msm_affine.cuh):
// array y lives at in + y * in_stride, out + y * out_stride, block_sums + y * bs_stride.
static __global__ void __launch_bounds__(1024) k_scan_blocks(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                      uint32_t *__restrict__ block_sums, uint32_t n, size_t in_stride = 0,
                                                      size_t out_stride = 0, size_t bs_stride = 0) {
    __shared__ uint32_t warp_tot[32];
    in += blockIdx.y * in_stride; out += blockIdx.y * out_stride; block_sums += blockIdx.y * bs_stride;
    uint32_t base = blockIdx.x * DG_SCAN_ITEMS + threadIdx.x * 4;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = warp_tot[lane], inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        warp_tot[lane] = inc - t;
        if (lane == 31) block_sums[blockIdx.x] = inc;
    }
    __syncthreads();
    uint32_t excl = warp_tot[wid] + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
}
static __global__ void __launch_bounds__(1024) k_scan_sums(uint32_t *block_sums, uint32_t nblocks, size_t bs_stride = 0) {
    // single block per array; serial over tiles of 1024
    __shared__ uint32_t warp_tot[32];
    block_sums += blockIdx.y * bs_stride;
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t t0 = 0; t0 < nblocks; t0 += 1024) {
        uint32_t i = t0 + threadIdx.x;
        uint32_t v = i < nblocks ? block_sums[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t t = warp_tot[lane], inc = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            warp_tot[lane] = inc - t;
        }
        __syncthreads();
        uint32_t excl = running + warp_tot[wid] + incl - v;
        if (i < nblocks) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) running = excl + v;
        __syncthreads();
    }
}
static __global__ void __launch_bounds__(1024) k_scan_add(uint32_t *__restrict__ out, const uint32_t *__restrict__ block_sums,
                                                   const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ copy,
                                                   size_t in_stride = 0, size_t out_stride = 0, size_t bs_stride = 0) {
    in += blockIdx.y * in_stride; out += blockIdx.y * out_stride; block_sums += blockIdx.y * bs_stride;
    uint32_t base = blockIdx.x * DG_SCAN_ITEMS + threadIdx.x * 4, add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (base + k < n) {
            uint32_t v = out[base + k] + add;
            out[base + k] = v;
            if (copy) copy[base + k] = v;
            if (base + k == n - 1) out[n] = v + in[n - 1];
        }
}

// ---------------------------------------------------------------- bucket accumulation --------
// entries[] is sorted by bucket; off[b] .. off[b+1] is bucket b's slice; M = off[nb].
// Thread t owns entries [t*L, (t+1)*L).  Runs (maximal same-bucket stretches inside the chunk):
//   first run of the chunk  -> head[t]       last run (if not also first) -> tail[t]
//   runs strictly inside    -> buckets[b] directly (nobody else touches that bucket)
// ---- per-thread TMA staging of the gathered points ---------------------------------------------
// The base of the NEXT entry is fetched by a bulk asynchronous copy (cp.async.bulk, the TMA engine;
// UBLKCP in SASS) into a per-thread double buffer in shared memory while the current mixed addition
// runs, completion tracked by a per-thread mbarrier (transaction bytes).  The gather latency
// (L2 / HBM, the table of precomputed multiples does not fit L2) leaves the critical path and no
// registers are spent on the prefetched point.
__device__ __forceinline__ uint32_t dg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dg_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void dg_bulk_fetch(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    // order this thread's earlier generic-proxy reads of `dst` before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dg_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dg_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(dg_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dg_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(dg_smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}
#define DG_ACC_THREADS 128
template <class F> constexpr size_t dg_acc_smem_bytes() { return (2 * sizeof(Affine<F>) + 16) * DG_ACC_THREADS; }

// DIRECT: the points are the dense output of the batch-affine rounds (msm_affine.cuh): entry e IS
// point e, no sign.
template <class F, bool DIRECT>
__global__ void __launch_bounds__(DG_ACC_THREADS, (sizeof(F) > 48 ? 2 : 3)) k_accumulate(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ entries,
                                                       const uint32_t *__restrict__ off, uint32_t nb, uint32_t L,
                                                       XYZZ<F> *__restrict__ buckets, XYZZ<F> *__restrict__ head,
                                                       XYZZ<F> *__restrict__ tail) {
    extern __shared__ __align__(16) unsigned char dg_acc_smem[];
    constexpr uint32_t REC = sizeof(Affine<F>);
    unsigned char *stage[2] = {dg_acc_smem + (size_t)threadIdx.x * REC, dg_acc_smem + (size_t)(DG_ACC_THREADS + threadIdx.x) * REC};
    uint64_t *bars = reinterpret_cast<uint64_t *>(dg_acc_smem + 2 * (size_t)DG_ACC_THREADS * REC);
    uint64_t *bar[2] = {&bars[threadIdx.x], &bars[DG_ACC_THREADS + threadIdx.x]};

    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t M = off[nb];
    uint64_t start64 = (uint64_t)t * L;
    if (start64 >= M) return;
    uint32_t start = (uint32_t)start64;
    uint32_t end = (M - start > L) ? start + L : M;
    dg_mbar_init(bar[0], 1);
    dg_mbar_init(bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t IDX_MASK = DIRECT ? 0xffffffffu : 0x7fffffffu;
    uint32_t ent = DIRECT ? start : __ldg(&entries[start]);
    dg_bulk_fetch(stage[0], &bases[ent & IDX_MASK], REC, bar[0]);
    uint32_t ent1 = (start + 1 < end) ? (DIRECT ? start + 1 : __ldg(&entries[start + 1])) : 0;
    // bucket containing `start`: largest b with off[b] <= start  (and off[b+1] > start)
    uint32_t lo = 0, hi = nb;            // invariant off[lo] <= start < off[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t b = lo, bend = off[b + 1];
    XYZZ<F> acc = xyzz_inf<F>();
    bool first = true;
    for (uint32_t e = start; e < end; e++) {
        const uint32_t k = e - start, buf = k & 1;
        uint32_t ent2 = 0;
        if (e + 1 < end) {                                   // prefetch the next point into the other buffer
            dg_bulk_fetch(stage[buf ^ 1], &bases[ent1 & IDX_MASK], REC, bar[buf ^ 1]);
            if (e + 2 < end) ent2 = DIRECT ? e + 2 : __ldg(&entries[e + 2]);
        }
        if (e == bend) {
            if (first) xyzz_store(&head[t], acc); else xyzz_store(&buckets[b], acc);
            first = false;
            acc = xyzz_inf<F>();
            do { b++; bend = off[b + 1]; } while (bend <= e);
        }
        dg_mbar_wait(bar[buf], (k >> 1) & 1);
        Affine<F> p = {fload_rw<F>(stage[buf]), fload_rw<F>(stage[buf] + sizeof(F))};
        if (!DIRECT) p.y = fcneg(p.y, (ent >> 31) != 0);
        acc = xyzz_madd(acc, p);
        ent = ent1;
        ent1 = ent2;
    }
    if (first) xyzz_store(&head[t], acc); else xyzz_store(&tail[t], acc);
}

// Buckets whose pieces span more than DG_LONG_PIECES chunks (hot buckets: skewed scalars such as
// the 0/1-heavy witnesses of real circuits, or the sparsely populated top window) are queued for
// k_fixup_long_part / k_fixup_long_final instead of being summed by one thread.
#define DG_LONG_PIECES 24
template <class F>
__global__ void __launch_bounds__(128) k_bucket_fixup(const uint32_t *__restrict__ off, uint32_t nb, uint32_t L,
                                                      XYZZ<F> *__restrict__ buckets, const XYZZ<F> *__restrict__ head,
                                                      const XYZZ<F> *__restrict__ tail, uint32_t *__restrict__ long_count,
                                                      uint32_t *__restrict__ long_list) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t s = off[b], e = off[b + 1], M = off[nb];
    if (s == e) { xyzz_store(&buckets[b], xyzz_inf<F>()); return; }
    uint32_t t0 = s / L, t1 = (e - 1) / L;
    bool first0 = (s == t0 * L);
    if (t0 == t1) {
        uint64_t cend = (uint64_t)(t0 + 1) * L;
        bool last0 = (e == (cend < M ? (uint32_t)cend : M));
        if (first0) xyzz_store(&buckets[b], xyzz_load<F>(&head[t0]));
        else if (last0) xyzz_store(&buckets[b], xyzz_load<F>(&tail[t0]));
        return;                                     // middle run: already written by k_accumulate
    }
    if (t1 - t0 > DG_LONG_PIECES) {
        long_list[atomicAdd(long_count, 1u)] = b;
        return;
    }
    XYZZ<F> acc = first0 ? xyzz_load<F>(&head[t0]) : xyzz_load<F>(&tail[t0]);
    for (uint32_t t = t0 + 1; t <= t1; t++) acc = xyzz_add(acc, xyzz_load<F>(&head[t]));
    xyzz_store(&buckets[b], acc);
}

// ---------------------------------------------------------------- bucket reduction, low depth ---
// S_w = sum_b (b + 1) B[w][b]  over  N = 2^(c-1)  buckets per window without a long serial chain.
// Write b = h * 2^LO + l.  With the line sums  C_l = sum_h B[h][l]  (columns)  and  R_h = sum_l B[h][l]
// (rows), the subset sums  U_j = sum_{b : bit j of b set} B[b]  are sums over half of the C's
// (j < LO) or half of the R's (j >= LO), and
//        S = T + sum_j 2^j U_j,      T = sum_h R_h.
// Every sum is a strided accumulation by the quads of a CTA followed by a shared-memory tree, so
// the dependent-operation depth is ~ len/QP + log2 QP per stage (three stages + j doublings) instead of
// the 27 operations per level x 5..6 levels of a multi-level running-sum scheme (measured 0.69 ms at
// N = 2^15, all of it latency; this scheme 0.24 ms).
template <class F> struct RedGeom {
    static constexpr int QP = sizeof(F) > 48 ? 16 : 32;                     // quads per CTA (33 KB of workspace)
    static constexpr int threads(bool wide) { return wide ? QuadLanes<true>::cta_threads(QP) : QuadLanes<false>::cta_threads(QP); }
    static constexpr int THREADS = threads(QuadWide<F>::value);             // 128 (G1: 4 lanes per quad), 256 (G2: 12 lanes per quad)
};

// quad-cooperative: ACC of quad `qi` := sum of the items item(k), k in [0, count) (skipping k with bit `bit` clear when
// bit >= 0); the total is valid in quad 0 after the shared-memory tree.
template <class F, class ItemFn>
__device__ __forceinline__ void red_cta_sum_fn(QuadWS<F> *wsall, uint32_t count, int bit, ItemFn item, const QuadCtx &qc) {
    constexpr int QP = RedGeom<F>::QP;
    const uint32_t qi = qc.qi;
    QuadWS<F> &ws = wsall[qi];
    enum { ACC = DG_Q_ACC, ITEM = DG_Q_ITEM };
    if (qc.active) {
        quad_set_inf(ws, ACC, qc);
        for (uint32_t k = qi; k < count; k += QP) {
            if (bit >= 0 && !((k >> bit) & 1u)) continue;                 // quad-uniform
            quad_load(ws, ITEM, item(k), qc);
            quad_add(ws, ACC, ACC, ITEM, qc);
        }
    }
    __syncthreads();
    for (uint32_t s = QP / 2; s > 0; s >>= 1) {
        if (qc.active && qi < s) {
            quad_load(ws, ITEM, reinterpret_cast<const XYZZ<F> *>(&wsall[qi + s].v[4 * ACC]), qc);
            quad_add(ws, ACC, ACC, ITEM, qc);
        }
        __syncthreads();
    }
}
template <class F>
__device__ __forceinline__ void red_cta_sum(QuadWS<F> *wsall, const XYZZ<F> *base, uint32_t count, size_t stride, int bit,
                                            const QuadCtx &qc) {
    red_cta_sum_fn<F>(wsall, count, bit, [=](uint32_t k) { return base + (size_t)k * stride; }, qc);
}

// Hot buckets queued by k_bucket_fixup (skewed scalars such as the 0/1-heavy witnesses of real circuits, or a sparse top
// window): the pieces of bucket i are split over S_i = min(32, ceil(pieces / 256)) CTAs, each a strided quad sum + tree
// (k_fixup_long_part), and a second CTA sums the S_i partials (k_fixup_long_final).  A 13 000-point bucket (10 % ones
// among 2^18 witnesses) took 0.36 ms when one CTA of 128 threads summed it with thread-level additions.
#define DG_LONG_SPLIT 32
template <class F> __device__ __forceinline__ uint32_t long_split(uint32_t np) {
    uint32_t s = (np + 255) / 256;
    return s > DG_LONG_SPLIT ? DG_LONG_SPLIT : (s ? s : 1);
}
template <class F>
__global__ void __launch_bounds__(RedGeom<F>::THREADS) k_fixup_long_part(const uint32_t *__restrict__ off, uint32_t L,
                                                                         const XYZZ<F> *__restrict__ head, const XYZZ<F> *__restrict__ tail,
                                                                         const uint32_t *__restrict__ long_count,
                                                                         const uint32_t *__restrict__ long_list, XYZZ<F> *__restrict__ part) {
    extern __shared__ __align__(16) unsigned char dg_smem_quad[];
    QuadWS<F> *wsall = reinterpret_cast<QuadWS<F> *>(dg_smem_quad);
    QuadCtx qc = quad_ctx<QuadWide<F>::value>();
    const uint32_t cnt = *long_count;
    for (uint32_t idx = blockIdx.x; idx < cnt * DG_LONG_SPLIT; idx += gridDim.x) {       // CTA-uniform loop
        const uint32_t i = idx / DG_LONG_SPLIT, si = idx % DG_LONG_SPLIT;
        const uint32_t b = long_list[i];
        const uint32_t s = off[b], e = off[b + 1];
        const uint32_t t0 = s / L, t1 = (e - 1) / L, np = t1 - t0 + 1;
        const bool first0 = (s == t0 * L);
        const uint32_t S = long_split<F>(np);
        if (si >= S) continue;
        const uint32_t lo = (uint32_t)(((uint64_t)np * si) / S), hi = (uint32_t)(((uint64_t)np * (si + 1)) / S);
        red_cta_sum_fn<F>(wsall, hi - lo, -1, [=](uint32_t k) {
            const uint32_t j = lo + k;
            return (j == 0 && !first0) ? &tail[t0] : &head[t0 + j];
        }, qc);
        if (qc.active && qc.qi == 0) quad_store(wsall[0], DG_Q_ACC, &part[(size_t)i * DG_LONG_SPLIT + si], qc);
        __syncthreads();                                                                  // workspace is reused by the next item
    }
}
template <class F>
__global__ void __launch_bounds__(RedGeom<F>::THREADS) k_fixup_long_final(const uint32_t *__restrict__ off, uint32_t L,
                                                                          XYZZ<F> *__restrict__ buckets, const uint32_t *__restrict__ long_count,
                                                                          const uint32_t *__restrict__ long_list, const XYZZ<F> *__restrict__ part) {
    extern __shared__ __align__(16) unsigned char dg_smem_quad[];
    QuadWS<F> *wsall = reinterpret_cast<QuadWS<F> *>(dg_smem_quad);
    QuadCtx qc = quad_ctx<QuadWide<F>::value>();
    const uint32_t cnt = *long_count;
    for (uint32_t i = blockIdx.x; i < cnt; i += gridDim.x) {
        const uint32_t b = long_list[i];
        const uint32_t s = off[b], e = off[b + 1];
        const uint32_t np = (e - 1) / L - s / L + 1;
        red_cta_sum<F>(wsall, part + (size_t)i * DG_LONG_SPLIT, long_split<F>(np), 1, -1, qc);
        if (qc.active && qc.qi == 0) quad_store(wsall[0], DG_Q_ACC, &buckets[b], qc);
        __syncthreads();
    }
}

// stage A: line sums.  grid (2^HI rows + 2^LO columns, nwin); lines[w][0 .. 2^HI) = R, lines[w][2^HI ..) = C
template <class F, bool WIDE>
__global__ void __launch_bounds__(RedGeom<F>::threads(WIDE)) k_red_lines(const XYZZ<F> *__restrict__ buckets, uint32_t nbw, int LO, int HI,
                                                                   XYZZ<F> *__restrict__ lines, uint32_t line_stride) {
    extern __shared__ __align__(16) unsigned char dg_smem_quad[];
    QuadWS<F> *wsall = reinterpret_cast<QuadWS<F> *>(dg_smem_quad);
    QuadCtx qc = quad_ctx<WIDE>();
    const uint32_t w = blockIdx.y, line = blockIdx.x, nrows = 1u << HI, ncols = 1u << LO;
    const XYZZ<F> *bw = buckets + (size_t)w * nbw;
    if (line < nrows) red_cta_sum<F>(wsall, bw + (size_t)line * ncols, ncols, 1, -1, qc);            // row h: contiguous
    else red_cta_sum<F>(wsall, bw + (line - nrows), nrows, ncols, -1, qc);                            // column l: stride 2^LO
    if (qc.active && qc.qi == 0) quad_store(wsall[0], DG_Q_ACC, &lines[(size_t)w * line_stride + line], qc);
}

// stage B: V_j = 2^j U_j for j < LO + HI, V_{LO+HI} = T.  grid (LO + HI + 1, nwin)
template <class F>
__global__ void __launch_bounds__(RedGeom<F>::THREADS) k_red_subsets(const XYZZ<F> *__restrict__ lines, uint32_t line_stride, int LO, int HI,
                                                                     XYZZ<F> *__restrict__ vout, uint32_t v_stride) {
    extern __shared__ __align__(16) unsigned char dg_smem_quad[];
    QuadWS<F> *wsall = reinterpret_cast<QuadWS<F> *>(dg_smem_quad);
    QuadCtx qc = quad_ctx<QuadWide<F>::value>();
    const uint32_t w = blockIdx.y, nrows = 1u << HI, ncols = 1u << LO;
    const int j = (int)blockIdx.x;
    const XYZZ<F> *R = lines + (size_t)w * line_stride, *Cc = R + nrows;
    if (j < LO) red_cta_sum<F>(wsall, Cc, ncols, 1, j, qc);
    else if (j < LO + HI) red_cta_sum<F>(wsall, R, nrows, 1, j - LO, qc);
    else red_cta_sum<F>(wsall, R, nrows, 1, -1, qc);
    if (qc.active && qc.qi == 0) {
        QuadWS<F> &ws = wsall[0];
        if (j < LO + HI)
            for (int k = 0; k < j; k++)
                if (!fis_zero(ws.v[4 * DG_Q_ACC + 2])) quad_dbl(ws, DG_Q_ACC, DG_Q_ACC, qc);
        quad_store(ws, DG_Q_ACC, &vout[(size_t)w * v_stride + j], qc);
    }
}

// stage C: S_w = sum_j V_j.  grid (1, nwin)
template <class F>
__global__ void __launch_bounds__(RedGeom<F>::THREADS) k_red_final(const XYZZ<F> *__restrict__ v, uint32_t v_stride, int nv,
                                                                   XYZZ<F> *__restrict__ wsum, uint32_t wsum_stride) {
    extern __shared__ __align__(16) unsigned char dg_smem_quad[];
    QuadWS<F> *wsall = reinterpret_cast<QuadWS<F> *>(dg_smem_quad);
    QuadCtx qc = quad_ctx<QuadWide<F>::value>();
    const uint32_t w = blockIdx.y;
    red_cta_sum<F>(wsall, v + (size_t)w * v_stride, (uint32_t)nv, 1, -1, qc);
    if (qc.active && qc.qi == 0) quad_store(wsall[0], DG_Q_ACC, &wsum[(size_t)w * wsum_stride], qc);
}

// window sums S_w (one XYZZ per window at stride) -> sum_w 2^(c w) S_w, Horner from the top,
// by one quad (the only unavoidable serial chain of the MSM: c * (nwin - 1) doublings).
// Window-group split (msm_host.cuh): the group of the high windows doubles its sum `extra_dbl` = c * w0 more times and
// leaves it as XYZZ in `out_xyzz`; the group of the low windows adds that record (`addend`) to its own sum, ORs the other
// group's error word into its own and writes the Jacobian result.
template <class F>
__global__ void __launch_bounds__(32) k_window_combine(const XYZZ<F> *__restrict__ wsum, uint32_t stride, int nwin, int c, Jac<F> *out,
                                                       int extra_dbl = 0, XYZZ<F> *out_xyzz = nullptr, const XYZZ<F> *addend = nullptr,
                                                       const uint32_t *flag_in = nullptr, uint32_t *flag_out = nullptr) {
    extern __shared__ __align__(16) unsigned char dg_smem_quad[];
    QuadWS<F> &ws = reinterpret_cast<QuadWS<F> *>(dg_smem_quad)[0];
    QuadCtx qc = quad_ctx<QuadWide<F>::value>();
    if (!qc.active || qc.qi != 0 || blockIdx.x != 0) return;
    enum { ACC = DG_Q_ACC, ITEM = DG_Q_ITEM };
    quad_load(ws, ACC, &wsum[(size_t)(nwin - 1) * stride], qc);
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++)
            if (!fis_zero(ws.v[4 * ACC + 2])) quad_dbl(ws, ACC, ACC, qc);
        quad_load(ws, ITEM, &wsum[(size_t)w * stride], qc);
        quad_add(ws, ACC, ACC, ITEM, qc);
    }
    for (int k = 0; k < extra_dbl; k++)
        if (!fis_zero(ws.v[4 * ACC + 2])) quad_dbl(ws, ACC, ACC, qc);
    if (addend) {
        quad_load(ws, ITEM, addend, qc);
        quad_add(ws, ACC, ACC, ITEM, qc);
    }
    __syncwarp(qc.mask);
    if (threadIdx.x == 0) {
        XYZZ<F> r = {ws.v[4 * ACC], ws.v[4 * ACC + 1], ws.v[4 * ACC + 2], ws.v[4 * ACC + 3]};
        if (out_xyzz) xyzz_store(out_xyzz, r);
        else jac_store(out, xyzz_to_jac(r));
        if (flag_in && flag_out && *flag_in) atomicOr(flag_out, *flag_in);
    }
}

// GLV base expansion: out[i] = P_i, out[n + i] = phi(P_i) = (beta x_i, y_i); beta is an Fp constant for both groups
// (the identity record x = y = 0 maps to itself).
template <class F>
__global__ void __launch_bounds__(256) k_glv_expand(const Affine<F> *__restrict__ in, uint32_t n, Affine<F> *__restrict__ out, uint32_t phi_off) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = aff_load<F>(&in[i]);
    Fp beta;
#pragma unroll
    for (int k = 0; k < 12; k++) beta.l[k] = sizeof(F) > 48 ? DGC_GLV_BETA_G2[k] : DGC_GLV_BETA_G1[k];
    if (out != in) aff_store(&out[i], p);
    p.x = glv_mul_beta(p.x, beta);
    aff_store(&out[phi_off + i], p);
}

}  // namespace dg
