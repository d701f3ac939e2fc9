// Host-side launch sequence of the Pippenger pipeline (see msm_kernels.cuh); included by
// msm_g1.cu and msm_g2.cu which instantiate it for Fp and Fp2.
#pragma once
#include "common.cuh"
#include "msm_kernels.cuh"
#include "msm_affine.cuh"

namespace dg {

static inline int ceil_log2_sz(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }

// Window bits: about log2(n) - 4 so that an average bucket receives ~32 points per window,
// which keeps the bucket reduction (2 * 2^(c-1) full additions per window) near 10 % of the
// accumulation work.  nwin * c >= 256 so the top signed digit cannot overflow.
// A run over the digit positions [w0, w0 + wcnt) of every scalar (plain bases only); wcnt = 0: all of them.
struct MsmPart { int w0, wcnt; };

static inline MsmGeom msm_geometry(size_t n, MsmPre pre, bool fp2 = false, MsmPart part = MsmPart{0, 0}) {
    int c = pre.c ? pre.c : ctx().msm_window_override.load();
    const bool glv = !pre.c && ctx().tunable[4].load() == 0;
    if (c <= 0) {
        const int logn = ceil_log2_sz(n ? n : 1);
        if (glv) {
            // measured on B200 with the GLV split (tools/sweep_rounds.py, 2^16 .. 2^24): 8 windows of 16 bits win from 2^17
            // terms up to at least 2^24 (2^15 buckets per window keep the reduction small, 8 x 16 = 128 bits exactly);
            // below that the window-combination chain dominates and about log2(n) - 3 bits are best
            c = logn >= 17 ? 16 : logn - 3;
            // G2 (tools/sweep_g2.py): the bucket reduction costs 3x as much per bucket, so 10 windows of 13 bits beat 8 of 16
            // up to 2^18 terms (7.40 vs 7.98 ms at 2^18); from 2^19 the accumulation dominates again (2^20: 19.0 vs 20.9 ms)
            if (fp2 && (logn == 17 || logn == 18)) c = 13;
            if (c < 4) c = 4;
        } else {
            c = logn - 4;
            if (c < 4) c = 4;
            if (c > 20) c = 20;
            if (c >= 13 && c <= 19) c = 16;   // measured on B200 (tools/sweep_c.py): 16 x 16 = 256 bits exactly wins from 2^17 to 2^23
        }
    }
    MsmGeom g;
    g.c = c;
    // plain bases: GLV split (tunable 4 != 0 switches it off, for A/B measurements and tests of the unsplit path)
    g.glv = glv ? 1 : 0;
    g.ndig = g.glv ? glv_ndigits(c) : msm_ndigits(c);
    g.nwin = pre.c ? 1 : g.ndig;
    g.w0 = 0;
    if (!pre.c && part.wcnt > 0) { g.w0 = part.w0; g.nwin = part.wcnt; }
    g.nbw = 1u << (c - 1);
    g.nb = g.nbw * (uint32_t)g.nwin;
    g.row_stride = pre.c ? pre.row_stride : 0;
    g.fp2 = fp2 ? 1 : 0;
    g.phi_off = 0;
    return g;
}

#define DG_BA_MAX_ROUNDS 10
struct MsmLayout {
    MsmGeom g;
    uint32_t L, nchunks, red_stride;
    size_t o_hist, o_off, o_cursor, o_bsums, o_entries, o_buckets, o_head, o_tail, o_long, o_longpart, o_red[4], o_glv, total;
    // batch-affine pre-reduction (msm_affine.cuh): R rounds, round r turns <= mb[r] points into <= mb[r + 1]
    int R;
    uint64_t mb[DG_BA_MAX_ROUNDS + 1];
    uint32_t K[DG_BA_MAX_ROUNDS], ctas[DG_BA_MAX_ROUNDS];
    size_t o_cnt, cnt_stride, o_offr[DG_BA_MAX_ROUNDS], o_aff[2], o_pre, o_rbsums, rbs_stride;
};

// Rounds of batch-affine halving before the XYZZ accumulation: each round costs ~6.3 instead of 10
// multiplications per addition but has a fixed cost (launch, one inversion per CTA batch), so it
// pays while the buckets still hold several points each.
static inline int msm_affine_rounds(size_t n, const MsmGeom &g) {
    int ov = ctx().msm_rounds_override.load();
    if (ov >= 0) return ov > DG_BA_MAX_ROUNDS ? DG_BA_MAX_ROUNDS : ov;
    // measured (tools/sweep_rounds.py): a round pays while the buckets still hold >= 6 points and it
    // has >= 2^20 (G1) / 2^18 (G2) additions to spread over the grid
    double entries = (double)n * (g.row_stride ? g.ndig : g.nwin) * (g.glv ? 2 : 1), load = entries / (double)g.nb;     // average points per bucket
    int r = 0;
    const double min_adds = g.fp2 ? 262144.0 : 1048576.0;              // an Fp2 addition is ~3x the work: smaller rounds still pay
    while (r < DG_BA_MAX_ROUNDS && load >= 6.0 && entries * 0.5 >= min_adds) { load *= 0.5; entries *= 0.5; r++; }
    return r;
}

template <class F> static inline MsmLayout msm_layout(size_t n, MsmPre pre, MsmPart part = MsmPart{0, 0}) {
    MsmLayout m;
    m.g = msm_geometry(n, pre, sizeof(F) > 48, part);
    size_t max_entries = n * (size_t)(pre.c ? m.g.ndig : m.g.nwin) * (m.g.glv ? 2 : 1);
    m.R = msm_affine_rounds(n, m.g);
    m.mb[0] = max_entries;
    for (int r = 0; r < m.R; r++) m.mb[r + 1] = (m.mb[r] + m.g.nb) / 2 + 1;     // sum_b ceil(n_b / 2) <= (M + nb) / 2
    // K outputs per thread: whole resident waves of CTAs (a fractional last wave costs a full one), at most kmax
    // outputs per thread.  tunable 0 = minimum number of waves, tunable 3 = kmax override.
    int min_waves = ctx().tunable[0].load();
    if (min_waves < 1) min_waves = 1;
    size_t kmax = ctx().tunable[3].load() > 0 ? (size_t)ctx().tunable[3].load() : 128;
    const size_t ba_threads = (size_t)ctx().sm_count * (sizeof(F) > 48 ? DG_BA_G2_CTAS : 4) * DG_BA_THREADS;     // one resident wave
    for (int r = 0; r < m.R; r++) {
        size_t waves = (m.mb[r + 1] + ba_threads * kmax - 1) / (ba_threads * kmax);
        if (waves < (size_t)min_waves) waves = min_waves;
        size_t k = (m.mb[r + 1] + ba_threads * waves - 1) / (ba_threads * waves);
        if (k < 16) k = 16;                     // one inversion per CTA batch of 128 x K additions
        m.K[r] = (uint32_t)k;
        m.ctas[r] = (uint32_t)((m.mb[r + 1] + k * DG_BA_THREADS - 1) / (k * DG_BA_THREADS));
    }
    const size_t acc_entries = m.mb[m.R];                                         // what k_accumulate folds
    size_t threads_target = (size_t)ctx().sm_count * 384 * 4;
    size_t L = (acc_entries + threads_target - 1) / threads_target;
    if (L < 8) L = 8;
    if (L > 512) L = 512;
    m.L = (uint32_t)L;
    m.nchunks = (uint32_t)((acc_entries + L - 1) / L);
    if (m.nchunks == 0) m.nchunks = 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += Arena::pad(bytes); return r; };
    m.o_hist = take(sizeof(uint32_t) * m.g.nb);
    m.o_off = take(sizeof(uint32_t) * ((size_t)m.g.nb + 1));
    m.o_cursor = take(sizeof(uint32_t) * m.g.nb);
    m.o_bsums = take(sizeof(uint32_t) * (m.g.nb / DG_SCAN_ITEMS + 2));
    m.o_entries = take(sizeof(uint32_t) * (max_entries ? max_entries : 1));
    m.o_buckets = take(sizeof(XYZZ<F>) * m.g.nb);
    m.o_head = take(sizeof(XYZZ<F>) * m.nchunks);
    m.o_tail = take(sizeof(XYZZ<F>) * m.nchunks);
    m.o_long = take(sizeof(uint32_t) * (m.nchunks / DG_LONG_PIECES + 64));     // [0] = count, list from [16]
    m.o_longpart = take(sizeof(XYZZ<F>) * (size_t)(m.nchunks / DG_LONG_PIECES + 1) * DG_LONG_SPLIT);
    {   // reduction scratch: [0] line sums (2^HI + 2^LO per window), [1] weighted subset sums (<= 32 per window), [2] window sums
        int LB = 0;
        while ((1u << LB) < m.g.nbw) LB++;
        m.red_stride = (1u << (LB - LB / 2)) + (1u << (LB / 2));
        m.o_red[0] = take(sizeof(XYZZ<F>) * (size_t)m.g.nwin * m.red_stride);
        m.o_red[1] = take(sizeof(XYZZ<F>) * (size_t)m.g.nwin * 32);
        m.o_red[2] = take(sizeof(XYZZ<F>) * (size_t)m.g.nwin);
        m.o_red[3] = m.o_red[2];
    }
    m.o_glv = 0;
    if (m.g.glv && !pre.phi_off) m.o_glv = take(sizeof(Affine<F>) * 2 * (n ? n : 1));     // [P | phi(P)] built per call
    m.o_cnt = m.cnt_stride = m.o_pre = m.o_rbsums = m.rbs_stride = 0;
    m.o_aff[0] = m.o_aff[1] = 0;
    if (m.R) {
        m.cnt_stride = Arena::pad(sizeof(uint32_t) * m.g.nb) / sizeof(uint32_t);
        m.o_cnt = take(sizeof(uint32_t) * m.cnt_stride * m.R);
        for (int r = 0; r < m.R; r++) m.o_offr[r] = take(sizeof(uint32_t) * ((size_t)m.g.nb + 1));      // equal sizes: constant stride
        m.rbs_stride = Arena::pad(sizeof(uint32_t) * (m.g.nb / DG_SCAN_ITEMS + 2)) / sizeof(uint32_t);
        m.o_rbsums = take(sizeof(uint32_t) * m.rbs_stride * m.R);
        m.o_aff[0] = take(sizeof(Affine<F>) * m.mb[1]);                           // A_1, A_3, A_5
        if (m.R > 1) m.o_aff[1] = take(sizeof(Affine<F>) * m.mb[2]);              // A_2, A_4, A_6
        size_t pre = 0;
        for (int r = 0; r < m.R; r++) {
            size_t b = (size_t)m.K[r] * m.ctas[r] * DG_BA_THREADS * sizeof(F);
            if (b > pre) pre = b;
        }
        m.o_pre = take(pre);
    }
    m.total = o;
    return m;
}

template <class F> __global__ void k_set_jac_inf(Jac<F> *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) jac_store(out, jac_inf<F>());
}

// How a window group hands its sum over (window-group split, see msm_run below)
template <class F> struct MsmJoin {
    int extra_dbl = 0;                    // high group: c * w0 more doublings after its Horner chain
    XYZZ<F> *out_xyzz = nullptr;          // high group: the sum stays here as XYZZ
    const XYZZ<F> *addend = nullptr;      // low group: the high group's sum, added before the result is written
    const uint32_t *flag_in = nullptr;    // low group: the high group's error word, ORed into err_flag
    cudaEvent_t wait_ev = nullptr;        // low group: fired once addend / flag_in are final
};

template <class F>
static int32_t msm_run_part(const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, char *scratch,
                            uint32_t *err_flag, cudaStream_t s, MsmPre pre, const MsmStage *stage, MsmPart part, const MsmJoin<F> &join,
                            bool prof) {
    // The flag reports THIS run: a stale bit from an earlier asynchronous call must not fail a valid one.
    DG_CUDA(cudaMemsetAsync(err_flag, 0, 4, s));
    MsmLayout m = msm_layout<F>(n, pre, part);
    if (m.g.glv) {
        if (pre.phi_off) {
            m.g.phi_off = pre.phi_off;
        } else {                                           // raw bases: build [P | phi(P)] in the scratch first
            Affine<F> *ex = (Affine<F> *)(scratch + m.o_glv);
            DG_LAUNCH(k_glv_expand<F>, div_up(n, 256), 256, 0, s, (const Affine<F> *)bases_dev, (uint32_t)n, ex, (uint32_t)n);
            bases_dev = ex;
            m.g.phi_off = (uint32_t)n;
        }
    }
    const MsmGeom g = m.g;
    uint32_t *hist = (uint32_t *)(scratch + m.o_hist), *off = (uint32_t *)(scratch + m.o_off);
    uint32_t *cursor = (uint32_t *)(scratch + m.o_cursor), *bsums = (uint32_t *)(scratch + m.o_bsums);
    uint32_t *entries = (uint32_t *)(scratch + m.o_entries);
    XYZZ<F> *buckets = (XYZZ<F> *)(scratch + m.o_buckets), *head = (XYZZ<F> *)(scratch + m.o_head);
    XYZZ<F> *tail = (XYZZ<F> *)(scratch + m.o_tail);
    XYZZ<F> *red[4];
    for (int k = 0; k < 4; k++) red[k] = (XYZZ<F> *)(scratch + m.o_red[k]);

    DG_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * g.nb, s));
    unsigned gd = div_up(n, 256);
    if (stage && stage->nchunks > 1) {
        for (int k = 0; k < stage->nchunks; k++) {          // count each chunk as soon as its copy has landed
            const size_t lo = stage->lo[k], cnt = stage->lo[k + 1] - lo;
            DG_CUDA(cudaStreamWaitEvent(s, stage->ev[k], 0));
            if (cnt) DG_LAUNCH(k_digits<0>, div_up(cnt, 256), 256, 0, s, (const uint32_t *)scalars_dev + 8 * lo, (uint32_t)cnt, g, hist, (uint32_t *)nullptr, err_flag);
        }
    } else {
        DG_LAUNCH(k_digits<0>, gd, 256, 0, s, (const uint32_t *)scalars_dev, (uint32_t)n, g, hist, (uint32_t *)nullptr, err_flag);
    }
    unsigned sb = div_up(g.nb, DG_SCAN_ITEMS);
    DG_LAUNCH(k_scan_blocks, sb, 1024, 0, s, hist, off, bsums, g.nb);
    DG_LAUNCH(k_scan_sums, 1, 1024, 0, s, bsums, sb);
    DG_LAUNCH(k_scan_add, sb, 1024, 0, s, off, bsums, hist, g.nb, cursor);
    DG_LAUNCH(k_digits<1>, gd, 256, 0, s, (const uint32_t *)scalars_dev, (uint32_t)n, g, cursor, entries, err_flag);

    // stage 4a: R rounds of batch-affine pairwise halving (msm_affine.cuh)
    const uint32_t *acc_off = off;
    const Affine<F> *acc_points = (const Affine<F> *)bases_dev;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (prof && ctx().prof_enabled.load()) {
        DG_CUDA(cudaEventCreate(&pe0));
        DG_CUDA(cudaEventCreate(&pe1));
    }
    if (m.R) {
        uint32_t *cnt = (uint32_t *)(scratch + m.o_cnt);
        DG_LAUNCH(k_round_counts, div_up(g.nb, 256), 256, 0, s, off, g.nb, m.R, cnt, m.cnt_stride);
        {                                                      // all R offset arrays in three launches (blockIdx.y = round)
            uint32_t *offr0 = (uint32_t *)(scratch + m.o_offr[0]), *bs = (uint32_t *)(scratch + m.o_rbsums);
            const size_t ostr = m.R > 1 ? (m.o_offr[1] - m.o_offr[0]) / sizeof(uint32_t) : 0, bstr = m.rbs_stride;
            DG_LAUNCH(k_scan_blocks, dim3(sb, m.R), 1024, 0, s, cnt, offr0, bs, g.nb, m.cnt_stride, ostr, bstr);
            DG_LAUNCH(k_scan_sums, dim3(1, m.R), 1024, 0, s, bs, sb, bstr);
            DG_LAUNCH(k_scan_add, dim3(sb, m.R), 1024, 0, s, offr0, bs, cnt, g.nb, (uint32_t *)nullptr, m.cnt_stride, ostr, bstr);
        }
        uint4 *pre_scratch = (uint4 *)(scratch + m.o_pre);
        for (int r = 0; r < m.R; r++) {
            const uint32_t *off_in = r ? (const uint32_t *)(scratch + m.o_offr[r - 1]) : off;
            const uint32_t *off_out = (const uint32_t *)(scratch + m.o_offr[r]);
            Affine<F> *dst = (Affine<F> *)(scratch + m.o_aff[r & 1]);
            const Affine<F> *src = r ? (const Affine<F> *)(scratch + m.o_aff[(r - 1) & 1]) : (const Affine<F> *)bases_dev;
            if (r == 0 && pe0) DG_CUDA(cudaEventRecord(pe0, s));
            {
                auto kg = k_affine_round<F, true>;
                auto kd = k_affine_round<F, false>;
                if (r == 0) DG_LAUNCH(kg, m.ctas[r], DG_BA_THREADS, 0, s, src, entries, off_in, off_out, g.nb, m.K[r], dst, pre_scratch);
                else DG_LAUNCH(kd, m.ctas[r], DG_BA_THREADS, 0, s, src, (const uint32_t *)nullptr, off_in, off_out, g.nb, m.K[r], dst, pre_scratch);
            }
            if (r == 0 && pe0) DG_CUDA(cudaEventRecord(pe1, s));
        }
        acc_off = (const uint32_t *)(scratch + m.o_offr[m.R - 1]);
        acc_points = (const Affine<F> *)(scratch + m.o_aff[(m.R - 1) & 1]);
    }
    // stage 4b: XYZZ accumulation of what is left
    {                                                      // the Fp2 staging buffers exceed the 48 KB default
        int32_t rc = smem_opt_in(k_accumulate<F, false>, dg_acc_smem_bytes<F>());
        if (!rc) rc = smem_opt_in(k_accumulate<F, true>, dg_acc_smem_bytes<F>());
        if (rc) return rc;
    }
    if (m.R) {
        auto kfn = k_accumulate<F, true>;
        DG_LAUNCH(kfn, div_up(m.nchunks, DG_ACC_THREADS), DG_ACC_THREADS, dg_acc_smem_bytes<F>(), s, acc_points,
                  (const uint32_t *)nullptr, acc_off, g.nb, m.L, buckets, head, tail);
    } else {
        if (pe0) DG_CUDA(cudaEventRecord(pe0, s));
        auto kfn = k_accumulate<F, false>;
        DG_LAUNCH(kfn, div_up(m.nchunks, DG_ACC_THREADS), DG_ACC_THREADS, dg_acc_smem_bytes<F>(), s, acc_points,
                  entries, acc_off, g.nb, m.L, buckets, head, tail);
        if (pe0) DG_CUDA(cudaEventRecord(pe1, s));
    }
    if (pe0) {
        std::lock_guard<std::mutex> lk(ctx().mu);
        ctx().prof_events.emplace_back(pe0, pe1);
    }
    uint32_t *long_count = (uint32_t *)(scratch + m.o_long), *long_list = long_count + 16;
    DG_CUDA(cudaMemsetAsync(long_count, 0, 64, s));
    DG_LAUNCH(k_bucket_fixup<F>, div_up(g.nb, 128), 128, 0, s, acc_off, g.nb, m.L, buckets, head, tail, long_count, long_list);
    {
        constexpr unsigned QPL = RedGeom<F>::QP;
        const size_t smem_l = sizeof(QuadWS<F>) * QPL;
        int32_t rc = smem_opt_in(k_fixup_long_part<F>, smem_l);
        if (!rc) rc = smem_opt_in(k_fixup_long_final<F>, smem_l);
        if (rc) return rc;
        XYZZ<F> *part = (XYZZ<F> *)(scratch + m.o_longpart);
        DG_LAUNCH(k_fixup_long_part<F>, 4 * ctx().sm_count, RedGeom<F>::THREADS, smem_l, s, acc_off, m.L, head, tail, long_count, long_list, part);
        DG_LAUNCH(k_fixup_long_final<F>, ctx().sm_count, RedGeom<F>::THREADS, smem_l, s, acc_off, m.L, buckets, long_count, long_list, part);
    }

    // bucket reduction: line sums -> weighted subset sums -> window sums (three shallow stages, msm_kernels.cuh)
    const XYZZ<F> *wsum = nullptr;
    uint32_t wsum_stride = 0;
    {
        int LB = 0;
        while ((1u << LB) < g.nbw) LB++;                                  // nbw = 2^LB
        const int LO = LB / 2, HI = LB - LO;
        const uint32_t nlines = (1u << HI) + (1u << LO), nv = (uint32_t)(LO + HI + 1);
        XYZZ<F> *lines = red[0], *vbuf = red[1], *ws_out = red[2];
        constexpr unsigned QP = RedGeom<F>::QP;
        const size_t smem = sizeof(QuadWS<F>) * QP;
        // G2: the line sums keep 4-lane quads.  Measured on B200 at 2^18 terms (tools/sweep_g2.py): 12-lane quads make this
        // stage slower both when the grid is several waves deep (raw bases: 7.45 -> 7.62 ms) and when every CTA is resident
        // at once (resident table: 7.63 -> 7.95 ms) -- the stage is bound by its ~2^16 additions, not by the chain length --
        // while the short-grid stages after it (subset sums, window sums, window combination, hot-bucket sums) are 2x faster
        // with them.  tunable 6 = 2 forces 12-lane quads here (A/B switch).
        constexpr bool W = QuadWide<F>::value;
        const bool wide_lines = W && ctx().tunable[6].load() == 2;
        int32_t rc = smem_opt_in(k_red_lines<F, false>, smem);
        if (!rc && W) rc = smem_opt_in(k_red_lines<F, W>, smem);
        if (!rc) rc = smem_opt_in(k_red_subsets<F>, smem);
        if (!rc) rc = smem_opt_in(k_red_final<F>, smem);
        if (rc) return rc;
        if (wide_lines) DG_LAUNCH((k_red_lines<F, W>), dim3(nlines, g.nwin), RedGeom<F>::threads(W), smem, s, buckets, g.nbw, LO, HI, lines, m.red_stride);
        else DG_LAUNCH((k_red_lines<F, false>), dim3(nlines, g.nwin), RedGeom<F>::threads(false), smem, s, buckets, g.nbw, LO, HI, lines, m.red_stride);
        DG_LAUNCH(k_red_subsets<F>, dim3(nv, g.nwin), RedGeom<F>::THREADS, smem, s, lines, m.red_stride, LO, HI, vbuf, 32u);
        DG_LAUNCH(k_red_final<F>, dim3(1, g.nwin), RedGeom<F>::THREADS, smem, s, vbuf, 32u, (int)nv, ws_out, 1u);
        wsum = ws_out;
        wsum_stride = 1;
    }
    if (join.wait_ev) DG_CUDA(cudaStreamWaitEvent(s, join.wait_ev, 0));
    DG_LAUNCH(k_window_combine<F>, 1, 32, sizeof(QuadWS<F>), s, wsum, wsum_stride, g.nwin, g.c, (Jac<F> *)out_jac_dev, join.extra_dbl,
              join.out_xyzz, join.addend, join.flag_in, err_flag);
    DG_CUDA(cudaGetLastError());
    return DG_OK;
}

// Window-group split of one MSM over plain bases (A/B switch, off by default).  The stages behind the batch-affine rounds
// (bucket fix-up, line sums, subset sums, the Horner chain over the windows) run on a handful of CTAs and are bound by the
// latency of dependent field multiplications: in a single stream they leave the GPU nearly idle for ~1.5 of the ~7 ms of a
// 2^20-term MSM.  With the split the high windows [ndig - h, ndig) go through the whole pipeline on a second stream of
// higher priority, so they finish first and their latency-bound tail (including the c * (ndig - h) extra doublings that
// put their sum in place) runs underneath the low windows' rounds; the low group's window combination adds the two sums.
// Both groups read the same scalars and the same [P | phi(P)] array; each has its own scratch and error word.
// Measured on B200 (tools/split_ab.py, raw G1 bases, h = ndig / 2): 2^18 3.03 -> 3.24 ms, 2^20 6.86 -> 7.30 ms,
// 2^22 22.14 -> 21.65 ms; G2 2^18 7.41 -> 7.70 ms.  Every batch-affine round is already sized to whole resident waves with
// one inversion per CTA batch, so two half-size pipelines pay the per-launch and per-batch fixed costs twice and that
// outweighs the hidden tail below 2^22 terms: the split stays behind tunable 2 (h >= 2 = split with h high windows).
template <class F> static inline bool msm_split_plan(size_t n, MsmPre pre, bool allow, MsmPart &hi, MsmPart &lo) {
    if (!allow || pre.c) return false;
    const int t = ctx().tunable[2].load();
    if (t < 2) return false;
    MsmGeom g = msm_geometry(n, pre, sizeof(F) > 48);
    if (!g.glv || g.ndig < 4) return false;
    int h = t;
    if (h > g.ndig - 1) h = g.ndig - 1;
    hi = MsmPart{g.ndig - h, h};
    lo = MsmPart{0, g.ndig - h};
    return true;
}

struct MsmSplitLayout { size_t o_glv, o_xyzz, o_flag, o_lo, o_hi, total; };
template <class F> static inline MsmSplitLayout msm_split_layout(size_t n, MsmPre pre, MsmPart hi, MsmPart lo) {
    MsmSplitLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += Arena::pad(bytes); return r; };
    L.o_glv = take(pre.phi_off ? 0 : sizeof(Affine<F>) * 2 * n);
    L.o_xyzz = take(sizeof(XYZZ<F>));
    L.o_flag = take(256);
    MsmPre pre2 = pre;
    if (!pre2.phi_off) pre2.phi_off = (uint32_t)n;          // both groups see expanded bases
    L.o_lo = take(msm_layout<F>(n, pre2, lo).total);
    L.o_hi = take(msm_layout<F>(n, pre2, hi).total);
    L.total = o;
    return L;
}

template <class F> static inline size_t msm_scratch_total(size_t n, MsmPre pre, bool allow_split) {
    MsmPart hi, lo;
    if (n && msm_split_plan<F>(n, pre, allow_split, hi, lo)) return msm_split_layout<F>(n, pre, hi, lo).total;
    return msm_layout<F>(n, pre).total;
}

template <class F>
static int32_t msm_run(const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, char *scratch,
                       uint32_t *err_flag, cudaStream_t s, MsmPre pre, const MsmStage *stage = nullptr, bool allow_split = true) {
    if (n == 0) {
        DG_CUDA(cudaMemsetAsync(err_flag, 0, 4, s));
        DG_LAUNCH(k_set_jac_inf<F>, 1, 32, 0, s, (Jac<F> *)out_jac_dev);
        return DG_OK;
    }
    if (n >= (1ull << 31)) return fail(DG_ERR_BAD_ARG, "msm: n must be < 2^31");
    {
        MsmGeom g0 = msm_geometry(n, pre, sizeof(F) > 48);
        if ((uint64_t)n * g0.ndig * (g0.glv ? 2 : 1) >= 0xffffffffull)
            return fail(DG_ERR_BAD_ARG, "msm: n * digits must fit 32-bit entry offsets (n up to ~2^27)");
        if (g0.glv && (uint64_t)n + (pre.phi_off ? pre.phi_off : n) >= (1ull << 31))
            return fail(DG_ERR_BAD_ARG, "msm: n must be < 2^30 for plain bases");
    }
    if (pre.c && (uint64_t)pre.row_stride * msm_ndigits(pre.c) >= (1ull << 31))
        return fail(DG_ERR_BAD_ARG, "msm: precomputed table too large for 31-bit point indices");
    MsmPart hi, lo;
    if (!msm_split_plan<F>(n, pre, allow_split, hi, lo))
        return msm_run_part<F>(bases_dev, scalars_dev, n, out_jac_dev, scratch, err_flag, s, pre, stage, MsmPart{0, 0}, MsmJoin<F>(), true);

    ThreadState &t = tls();
    if (!t.split_stream) {
        int lo_pri = 0, hi_pri = 0;
        DG_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        DG_CUDA(cudaStreamCreateWithPriority(&t.split_stream, cudaStreamNonBlocking, hi_pri));
        DG_CUDA(cudaEventCreateWithFlags(&t.split_ev[0], cudaEventDisableTiming));
        DG_CUDA(cudaEventCreateWithFlags(&t.split_ev[1], cudaEventDisableTiming));
    }
    cudaStream_t s2 = t.split_stream;
    const MsmSplitLayout L = msm_split_layout<F>(n, pre, hi, lo);
    MsmPre pre2 = pre;
    if (!pre.phi_off) {                                     // raw bases: one expansion serves both groups
        Affine<F> *ex = (Affine<F> *)(scratch + L.o_glv);
        DG_LAUNCH(k_glv_expand<F>, div_up(n, 256), 256, 0, s, (const Affine<F> *)bases_dev, (uint32_t)n, ex, (uint32_t)n);
        bases_dev = ex;
        pre2.phi_off = (uint32_t)n;
    }
    XYZZ<F> *hi_sum = (XYZZ<F> *)(scratch + L.o_xyzz);
    uint32_t *hi_flag = (uint32_t *)(scratch + L.o_flag);
    DG_CUDA(cudaEventRecord(t.split_ev[0], s));             // everything queued on s so far (inputs, expansion, scratch reuse)
    DG_CUDA(cudaStreamWaitEvent(s2, t.split_ev[0], 0));
    MsmJoin<F> jh, jl;
    jh.extra_dbl = msm_geometry(n, pre2, sizeof(F) > 48).c * hi.w0;
    jh.out_xyzz = hi_sum;
    int32_t rc = msm_run_part<F>(bases_dev, scalars_dev, n, nullptr, scratch + L.o_hi, hi_flag, s2, pre2, stage, hi, jh, false);
    if (rc) return rc;
    DG_CUDA(cudaEventRecord(t.split_ev[1], s2));
    jl.addend = hi_sum;
    jl.flag_in = hi_flag;
    jl.wait_ev = t.split_ev[1];
    return msm_run_part<F>(bases_dev, scalars_dev, n, out_jac_dev, scratch + L.o_lo, err_flag, s, pre2, stage, lo, jl, false);
}

}  // namespace dg
