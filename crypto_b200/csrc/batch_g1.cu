// G1 entry points of the batch group operations + shared table bookkeeping + field test hook.
#include "batch_host.cuh"
using namespace dg;

namespace dg {
__global__ void __launch_bounds__(128) k_dbg_fp_op(int op, const Fp *a, const Fp *b, uint32_t n, Fp *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x = fp_load(&a[i]), y = fp_load(&b[i]), r;
    switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_sqr(x); break;
        case 4: r = fp_neg(x); break;
        case 5: r = fp_inv(x); break;            // Fermat a^(p-2)
        case 6: r = fp_inv_binary(x); break;     // bit-serial binary extended Euclid
        case 7: r = fp_inv_pornin(x); break;     // Pornin's binary GCD with 31-bit inner rounds (the one the kernels use)
        default: r = fp_zero();
    }
    fp_store(&out[i], r);
}
}  // namespace dg

namespace dg {
int32_t bases_precompute_g1(HandleRec &rec, int c, cudaStream_t s) { return bases_precompute<Fp>(rec, c, s); }
int32_t fold_ptrs_g1(const PtrList &pl, int k, void *out_jac_dev, cudaStream_t s) {
    DG_LAUNCH(k_fold_jac_ptrs<Fp>, 1, 32, 0, s, pl, (uint32_t)k, (Jac<Fp> *)out_jac_dev);
    DG_CUDA(cudaGetLastError());
    return DG_OK;
}
}

extern "C" {
int32_t dg_fixed_base_table_g1(const uint8_t *p, size_t hint_n, uint64_t *h) { return fixed_table_build<Fp>(p, hint_n, h); }
int32_t dg_fixed_base_mul_many_g1(uint64_t h, const uint8_t *s, size_t m, uint8_t *o) { return fixed_mul_many<Fp>(h, s, m, o); }
int32_t dg_fixed_base_mul_many_normalized_g1(uint64_t h, const uint8_t *s, size_t m, uint8_t *o) { return fixed_mul_many_normalized<Fp>(h, s, m, o); }
int32_t dg_batch_mul_g1(const uint8_t *p, const uint8_t *s, size_t m, uint8_t *o) { return batch_mul<Fp>(p, s, m, o); }
int32_t dg_batch_mul_add_fixed_g1(const uint8_t *p, const uint8_t *sa, uint64_t h, const uint8_t *sb, size_t m, uint8_t *o) {
    return batch_mul_add_fixed<Fp>(p, sa, h, sb, m, o);
}
int32_t dg_batch_mul_add_same_g1(const uint8_t *p, const uint8_t *sa, const uint8_t *v, const uint8_t *sb, size_t m, uint8_t *o) {
    if (!v) return fail(DG_ERR_BAD_ARG, "batch_mul_add_same: null pointer");
    return batch_mul_add_fixed<Fp>(p, sa, 0, sb, m, o, v);
}
int32_t dg_compress_g1(const uint8_t *l, const uint8_t *r, size_t m, const uint8_t *s, uint8_t *o) { return compress_host<Fp>(l, r, m, s, o); }
int32_t dg_normalize_batch_g1(const uint8_t *j, size_t m, uint8_t *o) { return normalize_host<Fp>(j, m, o); }
int32_t dg_fold_g1(const uint8_t *j, size_t k, uint8_t *o) { return fold_host<Fp>(j, k, o); }
int32_t dg_fold_g1_device(const void *j, size_t k, void *o, void *stream) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!o || (k && !j)) return fail(DG_ERR_BAD_ARG, "fold_device: null pointer");
    cudaStream_t s = stream ? (cudaStream_t)stream : tls().stream;
    DG_LAUNCH(k_fold_jac<Fp>, 1, 32, 0, s, (const Jac<Fp> *)j, (uint32_t)k, (Jac<Fp> *)o);
    return DG_OK;
}

int32_t dg_fixed_base_table_info(uint64_t handle, int32_t *window, int32_t *num_windows, int32_t *is_g2) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || (it->second.kind != HandleRec::TABLE_G1 && it->second.kind != HandleRec::TABLE_G2))
        return fail(DG_ERR_BAD_ARG, "fixed_base_table_info: bad handle");
    if (window) *window = it->second.window;
    if (num_windows) *num_windows = it->second.nwin;
    if (is_g2) *is_g2 = it->second.kind == HandleRec::TABLE_G2;
    return DG_OK;
}
int32_t dg_fixed_base_table_download(uint64_t handle, uint8_t *out_affine) {
    int32_t rc = check_init();
    if (rc) return rc;
    HandleRec r;
    {
        std::lock_guard<std::mutex> lk(ctx().mu);
        auto it = ctx().handles.find(handle);
        if (it == ctx().handles.end() || (it->second.kind != HandleRec::TABLE_G1 && it->second.kind != HandleRec::TABLE_G2))
            return fail(DG_ERR_BAD_ARG, "fixed_base_table_download: bad handle");
        r = it->second;
    }
    if (!out_affine) return fail(DG_ERR_BAD_ARG, "fixed_base_table_download: null pointer");
    size_t rec = r.kind == HandleRec::TABLE_G2 ? 192 : 96;
    DG_CUDA(cudaMemcpy(out_affine, r.dev, rec * r.n, cudaMemcpyDeviceToHost));
    return DG_OK;
}
int32_t dg_fixed_base_table_free(uint64_t handle) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || (it->second.kind != HandleRec::TABLE_G1 && it->second.kind != HandleRec::TABLE_G2))
        return fail(DG_ERR_BAD_ARG, "fixed_base_table_free: bad handle");
    cudaDeviceSynchronize();
    cudaFree(it->second.dev);
    ctx().handles.erase(it);
    return DG_OK;
}

int32_t dg_dbg_fp_op(int32_t op, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!a || !b || !out || n == 0) return fail(DG_ERR_BAD_ARG, "dbg_fp_op: null pointer");
    ThreadState &t = tls();
    rc = t.arena.ensure(3 * Arena::pad(48 * n), t.stream);
    if (rc) return rc;
    Fp *d_a = t.arena.alloc<Fp>(n), *d_b = t.arena.alloc<Fp>(n), *d_o = t.arena.alloc<Fp>(n);
    DG_CUDA(cudaMemcpyAsync(d_a, a, 48 * n, cudaMemcpyHostToDevice, t.stream));
    DG_CUDA(cudaMemcpyAsync(d_b, b, 48 * n, cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_dbg_fp_op, div_up(n, 128), 128, 0, t.stream, op, d_a, d_b, (uint32_t)n, d_o);
    DG_CUDA(cudaMemcpyAsync(out, d_o, 48 * n, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}
}  // extern "C"
