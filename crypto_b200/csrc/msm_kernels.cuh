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
//   5 k_bucket_fixup    stitches the partial sums of buckets that straddle chunks
//   6 k_reduce_level    multi-level weighted running-sum  sum_b b*B[w][b]  per window
//   7 k_window_combine  Horner over windows, result as Jacobian (ark Projective layout)
#pragma once
#include "ec.cuh"
#include "quad.cuh"

namespace dg {

struct MsmGeom {
    int c;            // window bits
    int ndig;         // signed digits per scalar, ndig * c >= 256 so the top digit never overflows
    int nwin;         // bucket sets: ndig, or 1 when the bases carry precomputed 2^(c*k) multiples
    uint32_t nbw;     // buckets per set = 2^(c-1)
    uint32_t nb;      // total buckets = nwin * nbw
    uint32_t row_stride;   // precomputed bases: row k (= 2^(c*k) * P_i) starts at k * row_stride; 0 = plain bases
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
    const uint32_t mask = (1u << g.c) - 1, half = 1u << (g.c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < g.ndig; w++) {
        int bit = w * g.c;
        uint32_t raw = 0;
        if (bit < 256) {
            int word = bit >> 5, off = bit & 31;
            uint64_t two = ((uint64_t)s[word + 1] << 32) | s[word];
            raw = (uint32_t)(two >> off) & mask;
        }
        uint32_t d = raw + carry;
        bool neg = d > half;
        carry = neg ? 1u : 0u;
        uint32_t mag = neg ? (1u << g.c) - d : d;
        if (mag != 0) {
            // precomputed rows fold every digit position into ONE bucket set: digit w of scalar i
            // selects the point 2^(c*w) * P_i stored at w * row_stride + i
            uint32_t bucket = (g.row_stride ? 0u : w * g.nbw) + mag - 1;
            if (PASS == 0) {
                atomicAdd(&counters[bucket], 1u);
            } else {
                uint32_t pos = atomicAdd(&counters[bucket], 1u);
                entries[pos] = (i + (uint32_t)w * g.row_stride) | (neg ? 0x80000000u : 0u);
            }
        }
    }
    if (PASS == 0 && carry) atomicOr(err_flag, 1u);   // scalar >= 2^(nwin*c - 1): not a canonical Fr
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
// k_bucket_fixup_long instead of being summed by one thread.
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

// One CTA per queued bucket: 128 threads each sum a contiguous slice of the pieces, then a
// shared-memory tree combines the 128 partial sums.
template <class F>
__global__ void __launch_bounds__(128) k_bucket_fixup_long(const uint32_t *__restrict__ off, uint32_t L,
                                                           XYZZ<F> *__restrict__ buckets, const XYZZ<F> *__restrict__ head,
                                                           const XYZZ<F> *__restrict__ tail, const uint32_t *__restrict__ long_count,
                                                           const uint32_t *__restrict__ long_list) {
    extern __shared__ __align__(16) unsigned char dg_smem_long[];
    XYZZ<F> *sm = reinterpret_cast<XYZZ<F> *>(dg_smem_long);
    uint32_t k = threadIdx.x, cnt = *long_count;
    for (uint32_t i = blockIdx.x; i < cnt; i += gridDim.x) {
        uint32_t b = long_list[i];
        uint32_t s = off[b], e = off[b + 1];
        uint32_t t0 = s / L, t1 = (e - 1) / L, np = t1 - t0 + 1;
        bool first0 = (s == t0 * L);
        uint32_t per = (np + 127) / 128, lo = k * per, hi = lo + per < np ? lo + per : np;
        XYZZ<F> acc = xyzz_inf<F>();
        for (uint32_t j = lo; j < hi; j++) {
            const XYZZ<F> *src = (j == 0 && !first0) ? &tail[t0] : &head[t0 + j];
            acc = xyzz_add(acc, xyzz_load<F>(src));
        }
        sm[k] = acc;
        __syncthreads();
        for (uint32_t st = 64; st > 0; st >>= 1) {
            if (k < st) sm[k] = xyzz_add(sm[k], sm[k + st]);
            __syncthreads();
        }
        if (k == 0) xyzz_store(&buckets[b], sm[0]);
        __syncthreads();
    }
}

// ---------------------------------------------------------------- bucket reduction -----------
// One level of  S = sum_j (j+1) x_j + sum_j y_j  per window.  QUAD (w, q) - four lanes sharing
// every point operation, see quad.cuh - folds the g items x[q*g .. q*g+g) with a running sum:
//   A = sum (k+1) x_{qg+k},  R = sum x_{qg+k},  Y = sum y.
// Then  S = sum_q (A_q + Y_q) + sum_{q>=1} q * (g R_q):  the next level's y'_q = A_q + Y_q and
// x'_{q-1} = g * R_q (log2 g doublings).  Items past the end are the identity.
template <class F>
__global__ void __launch_bounds__(64) k_reduce_level(const XYZZ<F> *__restrict__ x, uint32_t cnt_x, uint32_t stride_x,
                                                     const XYZZ<F> *__restrict__ y, uint32_t cnt_y, uint32_t stride_y,
                                                     int log_g, uint32_t ngroups, int nwin,
                                                     XYZZ<F> *__restrict__ xo, XYZZ<F> *__restrict__ yo, uint32_t stride_o) {
    extern __shared__ __align__(16) unsigned char dg_smem_quad[];
    QuadWS<F> &ws = reinterpret_cast<QuadWS<F> *>(dg_smem_quad)[threadIdx.x >> 2];
    QuadCtx qc = quad_ctx();
    uint32_t gid = blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2);
    if (gid >= ngroups * (uint32_t)nwin) return;          // whole quad leaves together
    uint32_t w = gid / ngroups, q = gid % ngroups, g = 1u << log_g;
    const XYZZ<F> *xw = x + (size_t)w * stride_x;
    enum { RUN = 0, ACC = 1, ITEM = 2 };
    quad_set_inf(ws, RUN, qc);
    quad_set_inf(ws, ACC, qc);
    for (uint32_t k = g; k-- > 0;) {
        uint32_t j = q * g + k;
        if (j < cnt_x) {
            quad_load(ws, ITEM, &xw[j], qc);
            quad_add(ws, RUN, RUN, ITEM, qc);
        }
        quad_add(ws, ACC, ACC, RUN, qc);
    }
    if (cnt_y) {
        const XYZZ<F> *yw = y + (size_t)w * stride_y;
        for (uint32_t k = 0; k < g; k++) {
            uint32_t j = q * g + k;
            if (j < cnt_y) {
                quad_load(ws, ITEM, &yw[j], qc);
                quad_add(ws, ACC, ACC, ITEM, qc);
            }
        }
    }
    quad_store(ws, ACC, &yo[(size_t)w * stride_o + q], qc);
    if (q >= 1) {
        for (int k = 0; k < log_g; k++)
            if (!fis_zero(ws.v[4 * RUN + 2])) quad_dbl(ws, RUN, RUN, qc);
        quad_store(ws, RUN, &xo[(size_t)w * stride_o + q - 1], qc);
    }
}

// window sums S_w (one XYZZ per window at stride) -> sum_w 2^(c w) S_w, Horner from the top,
// by one quad (the only unavoidable serial chain of the MSM: c * (nwin - 1) doublings).
template <class F>
__global__ void __launch_bounds__(32) k_window_combine(const XYZZ<F> *__restrict__ wsum, uint32_t stride, int nwin, int c, Jac<F> *out) {
    extern __shared__ __align__(16) unsigned char dg_smem_quad[];
    QuadWS<F> &ws = reinterpret_cast<QuadWS<F> *>(dg_smem_quad)[0];
    if (threadIdx.x >= 4 || blockIdx.x != 0) return;
    QuadCtx qc = quad_ctx();
    enum { ACC = 1, ITEM = 2 };
    quad_load(ws, ACC, &wsum[(size_t)(nwin - 1) * stride], qc);
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++)
            if (!fis_zero(ws.v[4 * ACC + 2])) quad_dbl(ws, ACC, ACC, qc);
        quad_load(ws, ITEM, &wsum[(size_t)w * stride], qc);
        quad_add(ws, ACC, ACC, ITEM, qc);
    }
    if (threadIdx.x == 0) {
        XYZZ<F> r = {ws.v[4 * ACC], ws.v[4 * ACC + 1], ws.v[4 * ACC + 2], ws.v[4 * ACC + 3]};
        jac_store(out, xyzz_to_jac(r));
    }
}

}  // namespace dg
