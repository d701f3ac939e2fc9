// G2 instantiation of the Pippenger pipeline (VariableBaseMSM::msm_bigint for G2Projective,
// legogroth16/src/prover.rs:344 b_g2_query).
#include "msm_host.cuh"
namespace dg {
size_t msm_scratch_bytes_g2(size_t n, MsmPre pre, bool allow_split) { return msm_scratch_total<Fp2>(n, pre, allow_split); }
void msm_plan_g2(size_t n, MsmPre pre, int *c, int *rounds) {
    MsmLayout m = msm_layout<Fp2>(n, pre);
    *c = m.g.c;
    *rounds = m.R;
}
int32_t msm_run_g2(const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, char *scratch,
                   uint32_t *err_flag, cudaStream_t s, MsmPre pre, const MsmStage *stage, bool allow_split) {
    return msm_run<Fp2>(bases_dev, scalars_dev, n, out_jac_dev, scratch, err_flag, s, pre, stage, allow_split);
}
int32_t glv_expand_g2(const void *in, size_t n, void *out, size_t phi_off, cudaStream_t s) {
    if (n) DG_LAUNCH(k_glv_expand<Fp2>, div_up(n, 256), 256, 0, s, (const Affine<Fp2> *)in, (uint32_t)n, (Affine<Fp2> *)out, (uint32_t)phi_off);
    DG_CUDA(cudaGetLastError());
    return DG_OK;
}
}  // namespace dg
