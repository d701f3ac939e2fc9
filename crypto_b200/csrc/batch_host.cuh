// Host wrappers (staging + launches) for the batch group operations; instantiated for G1 in
// batch_g1.cu and G2 in batch_g2.cu.
#pragma once
#include "batch_kernels.cuh"

namespace dg {

static inline int ln_without_floats(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l * 69 / 100; }
// ark FixedBase::get_mul_window_size, the rule utils::msm::WindowTable::new applies (utils/src/msm.rs:20)
static inline int fixed_base_window(size_t n) { return n < 32 ? 3 : ln_without_floats(n); }

template <class F> struct Sizes { static constexpr size_t AFF = 2 * sizeof(F), JAC = 3 * sizeof(F); };

template <class F> static int32_t normalize_device(const Jac<F> *d_in, size_t m, Affine<F> *d_out, F *d_prefix, cudaStream_t s) {
    if (m == 0) return DG_OK;
    DG_LAUNCH(k_normalize<F>, div_up(div_up(m, DG_NORM_CHUNK), 128), 128, 0, s, d_in, (uint32_t)m, d_out, d_prefix);
    return DG_OK;
}

template <class F> static int32_t fixed_table_build(const uint8_t *point, size_t hint_n, uint64_t *handle) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!point || !handle) return fail(DG_ERR_BAD_ARG, "fixed_base_table: null pointer");
    ThreadState &t = tls();
    int window = fixed_base_window(hint_n);
    int outerc = (255 + window - 1) / window;
    size_t total = (size_t)outerc << window;
    size_t need = Arena::pad(Sizes<F>::AFF) + Arena::pad(Sizes<F>::JAC * outerc) + Arena::pad(Sizes<F>::JAC * total) +
                  Arena::pad(sizeof(F) * total);
    rc = t.arena.ensure(need, t.stream);
    if (rc) return rc;
    Affine<F> *d_g = t.arena.alloc<Affine<F>>(1);
    Jac<F> *d_outer = t.arena.alloc<Jac<F>>(outerc);
    Jac<F> *d_rows = t.arena.alloc<Jac<F>>(total);
    F *d_prefix = t.arena.alloc<F>(total);
    Affine<F> *d_table = nullptr;
    DG_CUDA(cudaMalloc(&d_table, Sizes<F>::AFF * total));
    cudaError_t e = cudaMemcpyAsync(d_g, point, Sizes<F>::AFF, cudaMemcpyHostToDevice, t.stream);
    if (e != cudaSuccess) { cudaFree(d_table); return fail(DG_ERR_CUDA, cudaGetErrorString(e)); }
    rc = smem_opt_in(k_fixed_outer_quad<F>, sizeof(QuadWS<F>));
    if (rc) { cudaFree(d_table); return rc; }
    DG_LAUNCH(k_fixed_outer_quad<F>, 1, 32, sizeof(QuadWS<F>), t.stream, d_g, window, outerc, d_outer);
    DG_LAUNCH(k_fixed_rows<F>, div_up(total, 128), 128, 0, t.stream, d_outer, window, outerc, d_rows);
    normalize_device<F>(d_rows, total, d_table, d_prefix, t.stream);
    e = cudaStreamSynchronize(t.stream);
    if (e != cudaSuccess) { cudaFree(d_table); return fail(DG_ERR_CUDA, cudaGetErrorString(e)); }
    std::lock_guard<std::mutex> lk(ctx().mu);
    uint64_t h = ctx().next_handle++;
    HandleRec r;
    r.kind = sizeof(F) == 48 ? HandleRec::TABLE_G1 : HandleRec::TABLE_G2;
    r.dev = d_table; r.n = total; r.window = window; r.nwin = outerc;
    ctx().handles[h] = r;
    *handle = h;
    return DG_OK;
}

template <class F> static int32_t lookup_table(uint64_t handle, HandleRec &out) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || it->second.kind != (sizeof(F) == 48 ? HandleRec::TABLE_G1 : HandleRec::TABLE_G2))
        return fail(DG_ERR_BAD_ARG, "bad fixed-base table handle");
    out = it->second;
    return DG_OK;
}

template <class F> static int32_t fixed_mul_many(uint64_t handle, const uint8_t *scalars, size_t m, uint8_t *out_jac) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (m && (!scalars || !out_jac)) return fail(DG_ERR_BAD_ARG, "fixed_base_mul_many: null pointer");
    HandleRec tb;
    rc = lookup_table<F>(handle, tb);
    if (rc) return rc;
    if (m == 0) return DG_OK;
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(32 * m) + Arena::pad(Sizes<F>::JAC * m), t.stream);
    if (rc) return rc;
    uint8_t *d_s = t.arena.alloc<uint8_t>(32 * m);
    Jac<F> *d_o = t.arena.alloc<Jac<F>>(m);
    DG_CUDA(cudaMemcpyAsync(d_s, scalars, 32 * m, cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_fixed_mul_many<F>, div_up(m, 128), 128, 0, t.stream, (const Affine<F> *)tb.dev, tb.window, tb.nwin, d_s,
              (uint32_t)m, d_o);
    DG_CUDA(cudaMemcpyAsync(out_jac, d_o, Sizes<F>::JAC * m, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

// FixedBase::msm followed by CurveGroup::normalize_batch without leaving the device
// (legogroth16/src/generator.rs:335-425, vb_accumulator/src/batch_utils.rs:498-509): m affine records.
template <class F> static int32_t fixed_mul_many_normalized(uint64_t handle, const uint8_t *scalars, size_t m, uint8_t *out_affine) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (m && (!scalars || !out_affine)) return fail(DG_ERR_BAD_ARG, "fixed_base_mul_many_normalized: null pointer");
    HandleRec tb;
    rc = lookup_table<F>(handle, tb);
    if (rc) return rc;
    if (m == 0) return DG_OK;
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(32 * m) + Arena::pad(Sizes<F>::JAC * m) + Arena::pad(Sizes<F>::AFF * m) + Arena::pad(sizeof(F) * m),
                        t.stream);
    if (rc) return rc;
    uint8_t *d_s = t.arena.alloc<uint8_t>(32 * m);
    Jac<F> *d_j = t.arena.alloc<Jac<F>>(m);
    Affine<F> *d_a = t.arena.alloc<Affine<F>>(m);
    F *d_prefix = t.arena.alloc<F>(m);
    DG_CUDA(cudaMemcpyAsync(d_s, scalars, 32 * m, cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_fixed_mul_many<F>, div_up(m, 128), 128, 0, t.stream, (const Affine<F> *)tb.dev, tb.window, tb.nwin, d_s,
              (uint32_t)m, d_j);
    normalize_device<F>(d_j, m, d_a, d_prefix, t.stream);
    DG_CUDA(cudaMemcpyAsync(out_affine, d_a, Sizes<F>::AFF * m, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

// Quad-cooperative chains (k_batch_mul_quad) halve the latency of a chain but spend ~1.6x the multiplications of the
// thread-per-element kernels (full XYZZ additions, idle lanes in the narrow waves), so they pay while the launch is
// latency-bound.  Measured on B200 (host calls, G1): batch_mul of 1 000 / 5 000 / 10 000 / 16 000 elements 1.01 / 1.39 /
// 1.99 / 2.78 ms against 2.18 / 2.32 / 2.44 / 2.64 ms with one thread per element; the fused update (two chains per
// element) 1.21 / 2.08 / 3.25 ms against 2.49 / 2.61 / 3.31 ms at 1 000 / 5 000 / 10 000 elements.
// tunable 5: 4 forces the quad kernels, 5 forbids them (6: two threads per element).
static inline bool batch_quad_pays(size_t m, int chains_per_elem, int force) {
    if (force == 4) return true;
    if (force == 5 || force == 6 || force == 1 || force == 2 || force == 3) return false;
    return m * chains_per_elem <= (size_t)ctx().sm_count * (chains_per_elem == 1 ? 96 : 128);     // ~14 000 elements / ~9 500 pairs
}
template <class F> static inline size_t batch_quad_scratch(size_t m, int chains_per_elem) {
    const size_t nch = m * chains_per_elem;
    return Arena::pad(sizeof(XYZZ<F>) * 16 * nch) + Arena::pad(sizeof(XYZZ<F>) * nch);
}
// out[i] = [sa_i] P_i (v == nullptr) or [sa_i] P_i + [sb_i] V, on the caller's stream; scratch comes from the arena
template <class F>
static int32_t batch_mul_quad_device(const Affine<F> *d_p, const uint8_t *d_sa, const Affine<F> *d_v, const uint8_t *d_sb, size_t m, Jac<F> *d_o,
                                     ThreadState &t) {
    const int per = d_v ? 2 : 1;
    const size_t nch = m * per;
    XYZZ<F> *d_tab = t.arena.alloc<XYZZ<F>>(16 * nch);
    XYZZ<F> *d_part = t.arena.alloc<XYZZ<F>>(nch);
    constexpr unsigned QP = BatchQuadGeom<F>::QP, TH = BatchQuadGeom<F>::THREADS;
    const size_t smem = sizeof(QuadWS<F>) * QP;
    int32_t rc = smem_opt_in(k_batch_mul_quad<F>, smem);
    if (!rc) rc = smem_opt_in(k_quad_finish<F>, smem);
    if (rc) return rc;
    DG_LAUNCH(k_batch_mul_quad<F>, div_up(nch, QP), TH, smem, t.stream, d_p, d_sa, d_v, d_sb, (uint32_t)nch, ctx().tunable[4].load() == 0 ? 1 : 0,
              d_tab, d_part);
    DG_LAUNCH(k_quad_finish<F>, div_up(m, QP), TH, smem, t.stream, (const XYZZ<F> *)d_part, (uint32_t)m, d_v ? 1 : 0, d_o);
    return DG_OK;
}

template <class F> static int32_t batch_mul(const uint8_t *points, const uint8_t *scalars, size_t m, uint8_t *out_jac) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (m && (!points || !scalars || !out_jac)) return fail(DG_ERR_BAD_ARG, "batch_mul: null pointer");
    if (m == 0) return DG_OK;
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(Sizes<F>::AFF * m) + Arena::pad(32 * m) + Arena::pad(Sizes<F>::JAC * m) + batch_quad_scratch<F>(m, 1), t.stream);
    if (rc) return rc;
    Affine<F> *d_p = t.arena.alloc<Affine<F>>(m);
    uint8_t *d_s = t.arena.alloc<uint8_t>(32 * m);
    Jac<F> *d_o = t.arena.alloc<Jac<F>>(m);
    DG_CUDA(cudaMemcpyAsync(d_p, points, Sizes<F>::AFF * m, cudaMemcpyHostToDevice, t.stream));
    DG_CUDA(cudaMemcpyAsync(d_s, scalars, 32 * m, cudaMemcpyHostToDevice, t.stream));
    const int force = ctx().tunable[5].load();
    if (batch_quad_pays(m, 1, force)) {
        rc = batch_mul_quad_device<F>(d_p, d_s, nullptr, nullptr, m, d_o, t);
        if (rc) return rc;
    } else {
        DG_LAUNCH(k_batch_mul<F>, div_up(m, 128), 128, 0, t.stream, d_p, d_s, (uint32_t)m, d_o, ctx().tunable[4].load() == 0 ? 1 : 0);
    }
    DG_CUDA(cudaMemcpyAsync(out_jac, d_o, Sizes<F>::JAC * m, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

// v_affine != nullptr: V given directly, no window table (always the joint form)
// out[i] = normalize(left[i] + [scalar] right[i]): utils::compress and Key::compress of the SnarkPack aggregation
// (legogroth16/src/aggregation/utils.rs:26-37, key.rs:118-143), one launch for the whole half-vector.
template <class F> static int32_t compress_host(const uint8_t *left, const uint8_t *right, size_t m, const uint8_t *scalar, uint8_t *out_affine) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (m && (!left || !right || !scalar || !out_affine)) return fail(DG_ERR_BAD_ARG, "compress: null pointer");
    if (m == 0) return DG_OK;
    ThreadState &t = tls();
    rc = t.arena.ensure(3 * Arena::pad(Sizes<F>::AFF * m) + Arena::pad(256) + Arena::pad(Sizes<F>::JAC * m) + Arena::pad(sizeof(F) * m), t.stream);
    if (rc) return rc;
    Affine<F> *d_l = t.arena.alloc<Affine<F>>(m), *d_r = t.arena.alloc<Affine<F>>(m), *d_a = t.arena.alloc<Affine<F>>(m);
    uint8_t *d_s = t.arena.alloc<uint8_t>(32);
    Jac<F> *d_o = t.arena.alloc<Jac<F>>(m);
    F *d_prefix = t.arena.alloc<F>(m);
    DG_CUDA(cudaMemcpyAsync(d_l, left, Sizes<F>::AFF * m, cudaMemcpyHostToDevice, t.stream));
    DG_CUDA(cudaMemcpyAsync(d_r, right, Sizes<F>::AFF * m, cudaMemcpyHostToDevice, t.stream));
    DG_CUDA(cudaMemcpyAsync(d_s, scalar, 32, cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_batch_mul<F>, div_up(m, 128), 128, 0, t.stream, d_r, d_s, (uint32_t)m, d_o, ctx().tunable[4].load() == 0 ? 1 : 0, 1,
              (const Affine<F> *)d_l);
    normalize_device<F>(d_o, m, d_a, d_prefix, t.stream);
    DG_CUDA(cudaMemcpyAsync(out_affine, d_a, Sizes<F>::AFF * m, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

template <class F>
static int32_t batch_mul_add_fixed(const uint8_t *points, const uint8_t *sa, uint64_t handle, const uint8_t *sb, size_t m,
                                   uint8_t *out_affine, const uint8_t *v_affine = nullptr) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (m && (!points || !sa || !sb || !out_affine)) return fail(DG_ERR_BAD_ARG, "batch_mul_add_fixed: null pointer");
    HandleRec tb;
    if (!v_affine) {
        rc = lookup_table<F>(handle, tb);
        if (rc) return rc;
    }
    if (m == 0) return DG_OK;
    ThreadState &t = tls();
    rc = t.arena.ensure(2 * Arena::pad(Sizes<F>::AFF * m) + 2 * Arena::pad(32 * m) + Arena::pad(Sizes<F>::JAC * m) +
                            Arena::pad(sizeof(F) * m) + Arena::pad(sizeof(JacZ<F>) * 8) + Arena::pad(Sizes<F>::AFF) + batch_quad_scratch<F>(m, 2), t.stream);
    if (rc) return rc;
    Affine<F> *d_v = t.arena.alloc<Affine<F>>(1);
    if (v_affine) DG_CUDA(cudaMemcpyAsync(d_v, v_affine, Sizes<F>::AFF, cudaMemcpyHostToDevice, t.stream));
    Affine<F> *d_p = t.arena.alloc<Affine<F>>(m);
    Affine<F> *d_a = t.arena.alloc<Affine<F>>(m);
    uint8_t *d_sa = t.arena.alloc<uint8_t>(32 * m), *d_sb = t.arena.alloc<uint8_t>(32 * m);
    Jac<F> *d_o = t.arena.alloc<Jac<F>>(m);
    F *d_prefix = t.arena.alloc<F>(m);
    JacZ<F> *d_vtbl = t.arena.alloc<JacZ<F>>(8);
    DG_CUDA(cudaMemcpyAsync(d_p, points, Sizes<F>::AFF * m, cudaMemcpyHostToDevice, t.stream));
    DG_CUDA(cudaMemcpyAsync(d_sa, sa, 32 * m, cudaMemcpyHostToDevice, t.stream));
    DG_CUDA(cudaMemcpyAsync(d_sb, sb, 32 * m, cudaMemcpyHostToDevice, t.stream));
    const int glv = ctx().tunable[4].load() == 0 ? 1 : 0;
    // Below ~2^16 elements the call is latency-bound (one thread per element cannot fill the GPU): both products on one
    // doubling chain.  Larger batches are throughput-bound and the window table's 29 mixed additions per element are
    // cheaper than 64 more full additions.  tunable 5: 1 forces the joint form, 2 the table form, 3 the one-thread-per-element joint kernel.
    const int force = ctx().tunable[5].load();
    const bool joint = v_affine || force == 1 || force == 3 || (force != 2 && m < 65536);
    if (joint && batch_quad_pays(m, 2, force)) {
        // small batches: one quad per product (see batch_quad_pays)
        rc = batch_mul_quad_device<F>(d_p, d_sa, v_affine ? d_v : (const Affine<F> *)tb.dev + 1, d_sb, m, d_o, t);
        if (rc) return rc;
    } else if (joint) {
        DG_LAUNCH(k_w4_table<F>, 1, 32, 0, t.stream, v_affine ? d_v : (const Affine<F> *)tb.dev + 1, d_vtbl);     // table[0][1] = V
        // below ~2^14 elements even the joint form leaves most of the GPU idle: two threads per element, shorter chains
        // (measured at 7 000 / 10 000 / 14 000 / 16 000 elements: quads 2.34 / 3.27 / 4.38 / 4.56 ms, two threads 2.71 / 3.33 / 3.97 / 4.01,
        //  one joint chain 3.58 / 3.63 / 3.73 / 3.78; FOUR threads per element -- every GLV half on its own chain -- 2.96 / 4.45 /
        //  4.94 / 5.34: the doubling chain is repeated four times and the launch turns throughput-bound; not kept)
        if ((m < 12288 || force == 1 || force == 6) && force != 3) DG_LAUNCH(k_batch_mul_add_split<F>, div_up(m, 64), 128, 0, t.stream, d_p, d_sa, d_vtbl, d_sb, (uint32_t)m, d_o, glv);
        else DG_LAUNCH(k_batch_mul_add_joint<F>, div_up(m, 128), 128, 0, t.stream, d_p, d_sa, d_vtbl, d_sb, (uint32_t)m, d_o, glv);
    } else {
        DG_LAUNCH(k_batch_mul_add_fixed<F>, div_up(m, 128), 128, 0, t.stream, d_p, d_sa, (const Affine<F> *)tb.dev, tb.window,
                  tb.nwin, d_sb, (uint32_t)m, d_o, glv);
    }
    normalize_device<F>(d_o, m, d_a, d_prefix, t.stream);
    DG_CUDA(cudaMemcpyAsync(out_affine, d_a, Sizes<F>::AFF * m, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

template <class F> static int32_t normalize_host(const uint8_t *jac, size_t m, uint8_t *out_affine) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (m && (!jac || !out_affine)) return fail(DG_ERR_BAD_ARG, "normalize_batch: null pointer");
    if (m == 0) return DG_OK;
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(Sizes<F>::JAC * m) + Arena::pad(Sizes<F>::AFF * m) + Arena::pad(sizeof(F) * m), t.stream);
    if (rc) return rc;
    Jac<F> *d_in = t.arena.alloc<Jac<F>>(m);
    Affine<F> *d_out = t.arena.alloc<Affine<F>>(m);
    F *d_prefix = t.arena.alloc<F>(m);
    DG_CUDA(cudaMemcpyAsync(d_in, jac, Sizes<F>::JAC * m, cudaMemcpyHostToDevice, t.stream));
    normalize_device<F>(d_in, m, d_out, d_prefix, t.stream);
    DG_CUDA(cudaMemcpyAsync(out_affine, d_out, Sizes<F>::AFF * m, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

// dg_bases_precompute: replace the resident bases by the table {2^(c*k) * P_i}, k < msm_ndigits(c) = ceil(254/c)
template <class F> static int32_t bases_precompute(HandleRec &rec, int c, cudaStream_t s) {
    if (rec.window) return fail(DG_ERR_BAD_ARG, "bases_precompute: handle already holds a precomputed table");
    if (c < 8 || c > 24) return fail(DG_ERR_BAD_ARG, "bases_precompute: window must be in [8, 24]");
    int rows = msm_ndigits(c);
    size_t n = rec.n;
    if ((uint64_t)n * rows >= (1ull << 31)) return fail(DG_ERR_BAD_ARG, "bases_precompute: table exceeds 2^31 points");
    Affine<F> *table = nullptr;
    Jac<F> *jac = nullptr;
    F *prefix = nullptr;
    size_t extra = n * (size_t)(rows - 1);
    DG_CUDA(cudaMalloc(&table, Sizes<F>::AFF * n * rows));
    cudaError_t e = cudaMalloc(&jac, Sizes<F>::JAC * extra);
    if (e == cudaSuccess) e = cudaMalloc(&prefix, sizeof(F) * extra);
    if (e != cudaSuccess) {
        cudaFree(table); cudaFree(jac); cudaFree(prefix);
        return fail(DG_ERR_OOM, std::string("bases_precompute: ") + cudaGetErrorString(e));
    }
    cudaMemcpyAsync(table, rec.dev, Sizes<F>::AFF * n, cudaMemcpyDeviceToDevice, s);
    DG_LAUNCH(k_precompute_rows<F>, div_up(n, 128), 128, 0, s, (const Affine<F> *)rec.dev, (uint32_t)n, c, rows, jac);
    normalize_device<F>(jac, extra, table + n, prefix, s);
    e = cudaStreamSynchronize(s);
    cudaFree(jac); cudaFree(prefix);
    if (e != cudaSuccess) { cudaFree(table); return fail(DG_ERR_CUDA, cudaGetErrorString(e)); }
    cudaFree(rec.dev);
    rec.dev = table;
    rec.window = c;
    rec.nwin = rows;
    rec.phi_off = 0;                                       // the table replaces the GLV-expanded plain array
    return DG_OK;
}

template <class F> static int32_t fold_host(const uint8_t *jac, size_t k, uint8_t *out_jac) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!out_jac || (k && !jac)) return fail(DG_ERR_BAD_ARG, "fold: null pointer");
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(Sizes<F>::JAC * (k + 1)), t.stream);
    if (rc) return rc;
    Jac<F> *d = t.arena.alloc<Jac<F>>(k + 1);
    if (k) DG_CUDA(cudaMemcpyAsync(d + 1, jac, Sizes<F>::JAC * k, cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_fold_jac<F>, 1, 32, 0, t.stream, d + 1, (uint32_t)k, d);
    DG_CUDA(cudaMemcpyAsync(out_jac, d, Sizes<F>::JAC, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

}  // namespace dg
