// G2 entry points of the batch group operations.
#include "batch_host.cuh"
using namespace dg;
namespace dg {
int32_t bases_precompute_g2(HandleRec &rec, int c, cudaStream_t s) { return bases_precompute<Fp2>(rec, c, s); }
int32_t fold_ptrs_g2(const PtrList &pl, int k, void *out_jac_dev, cudaStream_t s) {
    DG_LAUNCH(k_fold_jac_ptrs<Fp2>, 1, 32, 0, s, pl, (uint32_t)k, (Jac<Fp2> *)out_jac_dev);
    DG_CUDA(cudaGetLastError());
    return DG_OK;
}
}
extern "C" {
int32_t dg_fixed_base_table_g2(const uint8_t *p, size_t hint_n, uint64_t *h) { return fixed_table_build<Fp2>(p, hint_n, h); }
int32_t dg_fixed_base_mul_many_g2(uint64_t h, const uint8_t *s, size_t m, uint8_t *o) { return fixed_mul_many<Fp2>(h, s, m, o); }
int32_t dg_fixed_base_mul_many_normalized_g2(uint64_t h, const uint8_t *s, size_t m, uint8_t *o) { return fixed_mul_many_normalized<Fp2>(h, s, m, o); }
int32_t dg_batch_mul_g2(const uint8_t *p, const uint8_t *s, size_t m, uint8_t *o) { return batch_mul<Fp2>(p, s, m, o); }
int32_t dg_compress_g2(const uint8_t *l, const uint8_t *r, size_t m, const uint8_t *s, uint8_t *o) { return compress_host<Fp2>(l, r, m, s, o); }
int32_t dg_normalize_batch_g2(const uint8_t *j, size_t m, uint8_t *o) { return normalize_host<Fp2>(j, m, o); }
int32_t dg_fold_g2(const uint8_t *j, size_t k, uint8_t *o) { return fold_host<Fp2>(j, k, o); }
}
