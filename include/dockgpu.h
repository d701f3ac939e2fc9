/* dockgpu.h -- C ABI of the B200-native BLS12-381 MSM / fixed-base / multi-pairing backend.
 *
 * This is the drop-in boundary for docknetwork/crypto's hot path.  The
 * reference has no FFI for this path today: it calls the arkworks traits directly.  Each entry
 * point below names the reference call(s) a Rust binding would route to it; INTEGRATION.md shows
 * the `extern "C"` block and the ark-ec patch that does the routing.
 *
 * Conventions
 *  - Every function returns 0 (DG_OK) or a negative dg_status; no exceptions cross the boundary.
 *    dg_last_error() gives a thread-local human-readable message.
 *  - Caller owns all host memory; outputs go to caller-owned buffers.  The library owns device
 *    memory behind opaque 64-bit handles.
 *  - Field elements are little-endian Montgomery limbs exactly as ark-ff stores them
 *    (Fp: 48 B, R = 2^384; Fp2: c0||c1).  Points cross as packed records:
 *        G1 affine  96 B  x||y          G2 affine 192 B  x.c0||x.c1||y.c0||y.c1
 *        G1 Jacobian 144 B x||y||z      G2 Jacobian 288 B        (ark Projective{x,y,z})
 *        Fp12 576 B  c0.c0.c0 || c0.c0.c1 || c0.c1.c0 ... c1.c2.c1 (ark field order)
 *    The point at infinity is the all-zero affine record (x = y = 0 is not on either curve);
 *    Jacobian outputs use z = 0 (x = y = R, i.e. ark's Projective::zero()).
 *  - Scalars are canonical (non-Montgomery) 256-bit little-endian integers < r, i.e. the
 *    BigInt<4> that `Fr::into_bigint()` yields and `msm_bigint` receives; anything >= r is
 *    rejected with DG_ERR_BAD_ARG.
 *  - Thread safety: calls may be issued concurrently from many host threads (rayon workers);
 *    each call uses a per-thread stream and scratch arena.  dg_init is idempotent.
 *  - *_device variants take device pointers and a cudaStream_t (as void*) and do not
 *    synchronise; they exist so callers that already keep operands in HBM (proving keys,
 *    chained operations, benchmarks) avoid staging copies.
 */
#ifndef DOCKGPU_H
#define DOCKGPU_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DG_OK = 0,
    DG_ERR_BAD_ARG = -1,     /* null pointer, bad handle, scalar >= r ... */
    DG_ERR_CUDA = -2,        /* a CUDA runtime call failed (message in dg_last_error) */
    DG_ERR_OOM = -3,         /* device allocation failed */
    DG_ERR_NOT_INIT = -4
} dg_status;

/* ---- lifecycle -------------------------------------------------------------------------- */
/* Select the CUDA device this process drives and create the context.  device < 0 keeps the current
 * device.  Idempotent.  Same as dg_init_devices(&device, 1). */
int32_t dg_init(int32_t device);
/* One process driving several GPUs, as the reference is one process with rayon threads
 * (utils/src/macros.rs:68-84; SURVEY.md 8b "dg_init(devs, ndev)").  devices[0] is the primary device: every
 * single-GPU entry point below runs there and the *_sharded entry points combine their per-device partial
 * results there (peer access over NVLink is enabled from devices[0] to the others when available).  One
 * host worker thread per device is started on the first sharded call.  Calling it again with the same list
 * is a no-op; a different list needs dg_shutdown first. */
int32_t dg_init_devices(const int32_t *devices, int32_t ndev);
int32_t dg_device_count(int32_t *ndev);
int32_t dg_shutdown(void);
/* Copies the calling thread's last error message (NUL-terminated, truncated to cap). */
int32_t dg_last_error(char *buf, size_t cap);
/* Number of kernels this library has launched so far in this process (bench.py gpu_launches). */
uint64_t dg_launch_count(void);
/* Blocks until all work issued by the calling thread's stream has finished. */
int32_t dg_sync(void);

/* ---- resident bases (proving-key model: upload once, reuse across MSMs) -------------------
 * Replaces nothing in the reference API; it is the device-side cache for
 * legogroth16::ProvingKeyCommon vectors (legogroth16/src/data_structures.rs:151-168) and BBS+
 * SignatureParamsG1::h (bbs_plus/src/setup.rs:128-146). */
int32_t dg_bases_upload_g1(const uint8_t *affine, size_t n, uint64_t *handle);
int32_t dg_bases_upload_g2(const uint8_t *affine, size_t n, uint64_t *handle);
int32_t dg_bases_free(uint64_t handle);
/* Same, split into one contiguous base range per device of dg_init_devices (sizes differ by at most one):
 * shard d stays resident on devices[d].  The handle is accepted by dg_msm_*_sharded, dg_bases_precompute
 * (every device builds the table of its own range) and dg_bases_free. */
int32_t dg_bases_upload_g1_sharded(const uint8_t *affine, size_t n, uint64_t *handle);
int32_t dg_bases_upload_g2_sharded(const uint8_t *affine, size_t n, uint64_t *handle);
/* Optional, for bases that are reused across many MSMs: replaces the resident points by the table
 * { 2^(c*k) * P_i : k < ceil(254/c) } (ceil(254/c) x the memory, built once on the device).  MSMs
 * through this handle then fold every digit position into ONE bucket set: no window-combination
 * doublings and ceil(254/c) x fewer buckets to reduce.  c = 0 picks a default (17 below 2^22 points, 20 from there).
 * Results are identical group elements. */
int32_t dg_bases_precompute(uint64_t handle, int32_t window_bits);

/* ---- variable-base MSM ---------------------------------------------------------------------
 * ark_ec::VariableBaseMSM::msm_bigint(bases, bigints) for G1Projective / G2Projective
 * (legogroth16/src/prover.rs:215,286,299,363,404,456,592; via msm_unchecked at
 * bbs_plus/src/setup.rs:145,192, bbs_plus/src/proof.rs:580,
 * schnorr_pok/src/pok_generalized_pedersen.rs:97,153, vb_accumulator/src/batch_utils.rs:667,
 * vb_accumulator/src/witness.rs:415, utils/src/randomized_mult_checker.rs:100,
 * utils/src/pairs.rs:146,154).  The caller truncates to min(len) as arkworks does.
 * Exactly one of (bases_handle != 0, bases != NULL) selects the bases; with a handle, `n` may be
 * smaller than the uploaded count (prefix).  out: one Jacobian point. */
int32_t dg_msm_g1(uint64_t bases_handle, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac);
int32_t dg_msm_g2(uint64_t bases_handle, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac);
/* ark_ec::VariableBaseMSM::msm_unchecked(bases, &[Fr]): scalars arrive as Fr Montgomery limbs (R = 2^256,
 * the in-memory form of ark-ff) and are mapped to canonical integers on the device (into_bigint) before
 * the MSM -- the form bbs_plus/src/setup.rs:145 and schnorr_pok/src/pok_generalized_pedersen.rs:97 call. */
int32_t dg_msm_unchecked_g1(uint64_t bases_handle, const uint8_t *bases, const uint8_t *scalars_fr_mont, size_t n, uint8_t *out_jac);
int32_t dg_msm_unchecked_g2(uint64_t bases_handle, const uint8_t *bases, const uint8_t *scalars_fr_mont, size_t n, uint8_t *out_jac);
/* The same calls over every device of dg_init_devices from ONE host thread (SURVEY.md 8b "dg_msm_g1_sharded", 8e):
 * the bases are split by contiguous ranges (resident behind a *_sharded handle, or scattered from `bases` when the
 * handle is 0), each device's worker thread copies its slice of the scalars over its own PCIe link and runs the full
 * Pippenger pipeline, and devices[0] folds the per-device partial results with one kernel that reads them out of
 * the peers' memory over NVLink.  Same result as dg_msm_g1 on the whole input, bit for bit after normalisation.
 * Sharded calls are serialised against each other; they may run next to single-GPU calls from other threads. */
int32_t dg_msm_g1_sharded(uint64_t sharded_handle, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac);
int32_t dg_msm_g2_sharded(uint64_t sharded_handle, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac);
int32_t dg_msm_unchecked_g1_sharded(uint64_t sharded_handle, const uint8_t *bases, const uint8_t *scalars_fr_mont, size_t n, uint8_t *out_jac);
/* Fr::into_bigint for n elements (Montgomery -> canonical little-endian integers) */
int32_t dg_fr_into_bigint(const uint8_t *fr_mont, size_t n, uint8_t *out_canonical);
/* device-pointer variants; out_jac_dev is device memory (144 / 288 B) */
int32_t dg_msm_g1_device(const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, void *stream);
int32_t dg_msm_g2_device(const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, void *stream);
/* The device-pointer variants return as soon as the kernels are queued and therefore cannot report a scalar >= r:
 * such a scalar contributes nothing and raises a device-side flag owned by the calling thread.  dg_stream_status
 * synchronises `stream` (NULL: the thread's own stream) and returns DG_ERR_BAD_ARG if the most recent MSM this
 * thread queued saw one, DG_OK otherwise; the flag is cleared at the start of every MSM.  A thread's scratch arena
 * is shared by its calls: a call on a different stream first waits (on the device) for the previous one. */
int32_t dg_stream_status(void *stream);
/* resident (optionally precomputed) bases behind a handle, scalars and output in device memory */
int32_t dg_msm_g1_handle_device(uint64_t bases_handle, const void *scalars_dev, size_t n, void *out_jac_dev, void *stream);
int32_t dg_msm_g2_handle_device(uint64_t bases_handle, const void *scalars_dev, size_t n, void *out_jac_dev, void *stream);
/* Overrides the automatic window size (0 restores it); for tuning and tests. */
int32_t dg_msm_set_window(int32_t c);
/* Overrides the number of batch-affine halving rounds that precede the XYZZ bucket accumulation
 * (-1 restores the automatic choice, 0 disables the stage); for tuning and tests. */
int32_t dg_msm_set_affine_rounds(int32_t rounds);
/* The plan an n-term MSM would run with: window bits and batch-affine rounds (precomputed_window_bits
 * = the c given to dg_bases_precompute, 0 for plain bases). */
int32_t dg_msm_plan(size_t n, int32_t is_g2, int32_t precomputed_window_bits, int32_t *window_bits, int32_t *affine_rounds);

/* ---- fixed-base batch multiplication --------------------------------------------------------
 * utils::msm::WindowTable::new(num_multiplications, group_elem) (utils/src/msm.rs:18-30):
 * window = FixedBase::get_mul_window_size(hint_n), table[k][j] = j * 2^(k*window) * g. */
int32_t dg_fixed_base_table_g1(const uint8_t *point_affine, size_t hint_n, uint64_t *table_handle);
int32_t dg_fixed_base_table_g2(const uint8_t *point_affine, size_t hint_n, uint64_t *table_handle);
/* scalar_size / window_size / num_windows fields of WindowTable (utils/src/msm.rs:8-13) */
int32_t dg_fixed_base_table_info(uint64_t table_handle, int32_t *window, int32_t *num_windows, int32_t *is_g2);
/* Copies the affine table to the host, row-major num_windows x 2^window records (serialisation
 * of WindowTable, utils/src/msm.rs:7). */
int32_t dg_fixed_base_table_download(uint64_t table_handle, uint8_t *out_affine);
int32_t dg_fixed_base_table_free(uint64_t table_handle);
/* WindowTable::multiply_many / multiply_field_elems_with_same_group_elem (utils/src/msm.rs:38-62)
 * = FixedBase::msm: out m Jacobian points. */
int32_t dg_fixed_base_mul_many_g1(uint64_t table_handle, const uint8_t *scalars, size_t m, uint8_t *out_jac);
int32_t dg_fixed_base_mul_many_g2(uint64_t table_handle, const uint8_t *scalars, size_t m, uint8_t *out_jac);
/* FixedBase::msm + CurveGroup::normalize_batch fused on the device: m affine records
 * (legogroth16/src/generator.rs:335-425 CRS generation, "next" row f2; vb_accumulator/src/batch_utils.rs:498-509 Omega::new). */
int32_t dg_fixed_base_mul_many_normalized_g1(uint64_t table_handle, const uint8_t *scalars, size_t m, uint8_t *out_affine);
int32_t dg_fixed_base_mul_many_normalized_g2(uint64_t table_handle, const uint8_t *scalars, size_t m, uint8_t *out_affine);

/* ---- independent scalar multiplications ----------------------------------------------------
 * AffineRepr::mul_bigint inside cfg_iter! maps (vb_accumulator/src/witness.rs:190,229,278;
 * utils/src/randomized_pairing_check.rs:126,153,157): out[i] = [s_i] P_i, Jacobian.  Scalars are 256-bit integers
 * (BigInt<4>, not necessarily < r).  Canonical scalars use the GLV endomorphism, which like the MSM assumes points of
 * the prime-order subgroup (every point the reference passes: deserialisation validates). */
int32_t dg_batch_mul_g1(const uint8_t *points_affine, const uint8_t *scalars, size_t m, uint8_t *out_jac);
int32_t dg_batch_mul_g2(const uint8_t *points_affine, const uint8_t *scalars, size_t m, uint8_t *out_jac);
/* Fused accumulator witness update (vb_accumulator/src/witness.rs:269-284):
 * out[i] = normalize( [a_i] P_i + [b_i] V ), V given by its window table.  out: m affine. */
int32_t dg_batch_mul_add_fixed_g1(const uint8_t *points_affine, const uint8_t *scalars_a, uint64_t table_handle,
                                  const uint8_t *scalars_b, size_t m, uint8_t *out_affine);
/* Same result with V given as one affine point instead of its window table: for the batch sizes of the reference's
 * workloads (10^4 witnesses) building WindowTable::new(m, V) costs more than it saves on a GPU (255 sequential
 * doublings), so both products of an element share one doubling chain instead. */
int32_t dg_batch_mul_add_same_g1(const uint8_t *points_affine, const uint8_t *scalars_a, const uint8_t *v_affine,
                                 const uint8_t *scalars_b, size_t m, uint8_t *out_affine);

/* out[i] = normalize(left[i] + [scalar] right[i]), one scalar for the whole vector: utils::compress and Key::compress of
 * the SnarkPack GIPA rounds (legogroth16/src/aggregation/utils.rs:26-37, key.rs:118-143; "next" row f3). */
int32_t dg_compress_g1(const uint8_t *left_affine, const uint8_t *right_affine, size_t m, const uint8_t *scalar, uint8_t *out_affine);
int32_t dg_compress_g2(const uint8_t *left_affine, const uint8_t *right_affine, size_t m, const uint8_t *scalar, uint8_t *out_affine);

/* ---- CurveGroup::normalize_batch ------------------------------------------------------------
 * (vb_accumulator/src/witness.rs:193,232,284; batch_utils.rs:506,524,633,651). */
int32_t dg_normalize_batch_g1(const uint8_t *jac, size_t m, uint8_t *out_affine);
int32_t dg_normalize_batch_g2(const uint8_t *jac, size_t m, uint8_t *out_affine);

/* ---- pairing ---------------------------------------------------------------------------------
 * ark_ec::pairing::Pairing for Bls12_381 (bbs_plus/src/proof.rs:494,
 * legogroth16/src/verifier.rs:69-80, utils/src/randomized_pairing_check.rs:134,204-214):
 * multi_miller_loop drops pairs with an identity on either side; final_exponentiation returns
 * None (is_some = 0) iff the input is zero. */
int32_t dg_multi_miller_loop(const uint8_t *g1_affine, const uint8_t *g2_affine, size_t k, uint8_t *out_fp12);
int32_t dg_final_exponentiation(const uint8_t *in_fp12, uint8_t *out_fp12, int32_t *is_some);
int32_t dg_multi_pairing(const uint8_t *g1_affine, const uint8_t *g2_affine, size_t k, uint8_t *out_fp12);
/* nbatch independent pairing products in one call: product b covers counts[b] consecutive pairs of the
 * concatenated inputs; out_fp12 receives nbatch x 576 B.  The Miller loops share one launch and the
 * per-product tails (CTA product tree + final exponentiation) overlap on separate streams -- SnarkPack's
 * GIPA rounds issue six multi_pairings per round (legogroth16/src/aggregation/utils.rs:85-97). */
int32_t dg_multi_pairing_batch(const uint8_t *g1_affine, const uint8_t *g2_affine, const size_t *counts, size_t nbatch,
                               uint8_t *out_fp12);
/* result = 1 iff prod e(P_i, Q_i) == 1 */
int32_t dg_multi_pairing_is_one(const uint8_t *g1_affine, const uint8_t *g2_affine, size_t k, int32_t *result);
/* Target-group helpers used by RandomizedPairingChecker (right += out * m, left *= miller):
 * PairingOutput::mul_bigint and Fp12 multiplication. */
int32_t dg_gt_pow(const uint8_t *in_fp12, const uint8_t *scalar, uint8_t *out_fp12);
int32_t dg_fp12_mul(const uint8_t *a_fp12, const uint8_t *b_fp12, uint8_t *out_fp12);

/* ---- multi-GPU combine ------------------------------------------------------------------------
 * Folds k Jacobian partial results (one per rank, gathered by the host's NCCL all-gather) into
 * one: the "all-reduce under the group law" of SURVEY.md 8e. */
int32_t dg_fold_g1(const uint8_t *jac_points, size_t k, uint8_t *out_jac);
int32_t dg_fold_g1_device(const void *jac_points_dev, size_t k, void *out_jac_dev, void *stream);
int32_t dg_fold_g2(const uint8_t *jac_points, size_t k, uint8_t *out_jac);

/* ---- NTT over Fr ("next" row f1 of SURVEY.md 8f) ---------------------------------------------------
 * ark_poly::Radix2EvaluationDomain<Fr> of size 2^logn: fft_in_place / ifft_in_place (inverse = 1) and
 * the coset variants (coset = 1) with offset Fr::GENERATOR = 7, as legogroth16's witness map uses
 * them (legogroth16/src/r1cs_to_qap.rs:187-207).  Elements are Fr Montgomery limbs (4 x u64 LE,
 * R = 2^256) exactly as ark-ff stores them; natural order in and out; in place. */
int32_t dg_fr_ntt(uint8_t *data, uint32_t logn, int32_t inverse, int32_t coset);
/* device variant: data_dev and tmp_dev are 2^logn x 32 B device buffers (tmp is scratch) */
int32_t dg_fr_ntt_device(void *data_dev, void *tmp_dev, uint32_t logn, int32_t inverse, int32_t coset, void *stream);
/* Tail of LibsnarkReduction::witness_map_from_matrices (r1cs_to_qap.rs:187-207): a, b, c are the
 * 2^logn constraint evaluations; out_h = coset_ifft((coset_fft(ifft a) * coset_fft(ifft b) -
 * coset_fft(ifft c)) / Z(7)), the coefficients the h_query MSM consumes. */
int32_t dg_qap_h_from_abc(const uint8_t *a, const uint8_t *b, const uint8_t *c, uint32_t logn, uint8_t *out_h);

/* Sparse constraint matrix times assignment over Fr (CSR; Montgomery elements in and out): the evaluate_constraint
 * map in front of the witness-map tail, LibsnarkReduction::witness_map_from_matrices
 * (legogroth16/src/r1cs_to_qap.rs:150-186).  out[i] = sum_{k in [row_ptr[i], row_ptr[i+1])} coeff[k] * w[col[k]]. */
int32_t dg_fr_spmv(const uint32_t *row_ptr, const uint32_t *col, const uint8_t *coeff_mont, size_t rows, size_t nnz,
                   const uint8_t *w_mont, size_t ncols, uint8_t *out_mont);

/* ---- device-chained LegoGroth16 prover (rows a18 + f1) ---------------------------------------------------------
 * The constraint matrices of a circuit (ark_relations ConstraintMatrices a, b, c as CSR with Fr Montgomery
 * coefficients; crypto_b200/r1cs.py produces this layout from a Circom .r1cs file) are uploaded once, next to the
 * proving key's bases.  row_ptr / col / coeff_mont each point at three arrays (a, b, c): row_ptr[k] has
 * num_constraints + 1 entries, col[k] / coeff_mont[k] have row_ptr[k][num_constraints] entries. */
int32_t dg_r1cs_upload(const uint32_t *const row_ptr[3], const uint32_t *const col[3], const uint8_t *const coeff_mont[3],
                       size_t num_constraints, size_t num_inputs, size_t num_vars, uint64_t *r1cs_handle);
int32_t dg_r1cs_free(uint64_t r1cs_handle);
/* create_proof_and_committed_witnesses_with_assignment's heavy half (legogroth16/src/prover.rs:267-383) for one full
 * assignment (instance then witness variables, Fr Montgomery, full_assignment[0] = 1), chained on the device:
 *   h = LibsnarkReduction::witness_map_from_matrices(...)           (r1cs_to_qap.rs:150-210)
 *   out_h_acc_jac = msm_bigint(h_query, h.into_bigint())             (prover.rs:281-286); h never visits the host
 *   out_jobs_jac[j] = msm_bigint(job_bases[j], assignment.into_bigint()[job_offset[j] .. job_offset[j] + job_count[j]])
 *                     for the l_query / a_query / b_g1_query / b_g2_query / gamma_abc MSMs (prover.rs:299,326,334,344,363);
 *                     each result sits in a 288-byte slot (G1 results use the first 144 bytes).
 * job_bases are dg_bases_upload_g1 / _g2 handles (optionally precomputed).  out_h_mont (may be NULL) receives the
 * 2^k coefficients of h for callers that want them. */
int32_t dg_groth16_prove_msms(uint64_t r1cs_handle, const uint8_t *full_assignment_mont, size_t num_vars, uint64_t h_query_handle,
                              const uint64_t *job_bases, const uint64_t *job_offset, const uint64_t *job_count, size_t njobs,
                              uint8_t *out_h_acc_jac, uint8_t *out_jobs_jac, uint8_t *out_h_mont);

/* ---- ark-serialize wire formats ("next" row f4 of SURVEY.md 8f) ------------------------------------
 * CanonicalSerialize::serialize_compressed / serialize_uncompressed and CanonicalDeserialize::
 * deserialize_compressed / deserialize_uncompressed for vectors of BLS12-381 points, as the reference
 * reaches them through ArkObjectBytes (utils/src/serde_utils.rs:13-33) and the derives on the proving /
 * verifying keys (legogroth16/src/data_structures.rs:7-189).  Encoded records are 48 / 96 B (G1
 * compressed / uncompressed) and 96 / 192 B (G2), big-endian, Zcash flag bits, c1 before c0; the affine
 * side is the packed Montgomery record of this ABI.  validate != 0 adds the subgroup check of
 * Validate::Yes.  status (n bytes, may be NULL): 0 ok, 1 malformed (compression flag, coordinate >= p,
 * stray bits), 2 not on the curve, 3 not in the prime-order subgroup; rejected elements come back as the
 * identity record and are counted in *invalid_count (may be NULL) -- ark returns Err for the whole
 * vector when the count is non-zero.  Stricter than ark-bls12-381 0.4 (as recalled; see csrc/serialize.cu) on two kinds
 * of malformed input no serializer emits: an infinity encoding with a non-zero body is status 1, and an uncompressed
 * point off the curve is status 2 even with validate == 0. */
int32_t dg_g1_serialize(const uint8_t *affine, size_t n, int32_t compressed, uint8_t *out);
int32_t dg_g2_serialize(const uint8_t *affine, size_t n, int32_t compressed, uint8_t *out);
int32_t dg_g1_deserialize(const uint8_t *in, size_t n, int32_t compressed, int32_t validate, uint8_t *out_affine, uint8_t *status,
                          size_t *invalid_count);
int32_t dg_g2_deserialize(const uint8_t *in, size_t n, int32_t compressed, int32_t validate, uint8_t *out_affine, uint8_t *status,
                          size_t *invalid_count);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------------
 * While enabled, every MSM records a CUDA-event pair on its launching stream around the bucket
 * accumulation kernel (the dominant kernel); dg_prof_read_accumulate synchronises the device,
 * returns the mean duration of the launches recorded since the last read, and clears the list. */
int32_t dg_prof_enable(int32_t on);
int32_t dg_prof_read_accumulate(double *mean_ms, int32_t *count);

/* ---- test hooks (field arithmetic parity; not part of the reference-facing surface) ---------- */
int32_t dg_dbg_fp_op(int32_t op, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out);
/* internal A/B switches for sweeps and tests (0 restores the default of each):
 *   0 minimum waves per batch-affine round      3 max outputs per thread of a batch-affine round
 *   1 streams of dg_groth16_prove_msms (2..5)   4 non-zero: no GLV split (MSM, batch multiplication)
 *   2 window-group split of an MSM over plain bases (high windows on a second stream, off by default): h >= 2 = h high windows
 *   5 form of the batch multiplications: 1, 6 two threads per element, 2 window table, 3 one joint chain per thread,
 *     4 one quad per product, 5 never quads
 *   6 2: 12-lane quads for the G2 line sums     7 non-zero: no chunked scalar staging in the host MSM path */
int32_t dg_dbg_set_tunable(int32_t id, int32_t value);
int32_t dg_dbg_fr_op(int32_t op, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* DOCKGPU_H */
