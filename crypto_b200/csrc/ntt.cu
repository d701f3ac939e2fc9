// Number-theoretic transform over the BLS12-381 scalar field on sm_100a -- SURVEY.md 8f row f1,
// the first "next" row after the MSM / pairing hot path.
//
// Replaces ark_poly::Radix2EvaluationDomain<Fr>::{fft_in_place, ifft_in_place} and the coset
// variants with offset Fr::GENERATOR = 7, and the tail of
// LibsnarkReduction::witness_map_from_matrices (legogroth16/src/r1cs_to_qap.rs:187-207):
// 3 iFFT + 3 coset FFT + pointwise (a*b - c)/Z(7) + 1 coset iFFT at D = 2^18..2^19, whose output
// `h` feeds the h_query MSM (legogroth16/src/prover.rs:286) without leaving the device.
//
// Layout: n = 2^k Fr elements, 32 B each (Montgomery limbs exactly as ark-ff stores them), natural
// order in and out.  The whole vector (8-16 MB at 2^18-2^19) lives in L2; the transform is a
// bit-reversal pass followed by radix-2 DIT stages grouped so that each CTA keeps a tile in shared
// memory for up to 8 (first pass, contiguous) or 6 (later passes, strided tiles of 8 x 32 B
// contiguous elements) stages: 3 passes over the data at 2^19.  Twiddles, coset powers and 1/n
// come from per-size tables built once on the device (plan cache).
#include <cstring>
#include <map>
#include "common.cuh"
#include "fr.cuh"

namespace dg {

struct NttPlan {
    uint32_t logn = 0;
    Fr *tw_fwd = nullptr, *tw_inv = nullptr;       // g^i, g^-i            (n/2 each)
    Fr *scale_fwd = nullptr, *scale_inv = nullptr; // 7^i ; n^-1 * 7^-i     (n each)
    Fr *consts = nullptr;                          // [0] n^-1  [1] 1/(7^n - 1)
};
static std::map<uint32_t, NttPlan> &plans() {
    static std::map<uint32_t, NttPlan> p;
    return p;
}

// consts layout produced by k_ntt_consts: 0 g, 1 g^-1, 2 n^-1, 3 7, 4 7^-1, 5 1/(7^n - 1)
__device__ Fr fr_pow_limbs(const Fr &a, const uint32_t *e, int nbits) {
    Fr acc = fr_one();
    for (int i = nbits - 1; i >= 0; i--) {
        acc = fr_mul(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fr_mul(acc, a);
    }
    return acc;
}
__device__ Fr fr_inv(const Fr &a) {                  // a^(r-2)
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = fr_mod_limb(i);
    e[0] = 0xffffffffu;                              // r - 2: r ends in ...ffffffff 00000001, the -2 borrows from limb 1
    e[1] -= 1;
    return fr_pow_limbs(a, e, 255);
}
__global__ void k_ntt_consts(uint32_t logn, Fr *c) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr g = fr_const(DGC_FR_W32);
    for (uint32_t i = logn; i < 32; i++) g = fr_mul(g, g);     // W32^(2^(32-logn))
    Fr gen = fr_const(DGC_FR_GEN);
    Fr nn = fr_zero();
    nn.l[0] = 1u << logn;                                        // logn <= 28
    nn = fr_mul(nn, fr_const(DGC_FR_R2));                        // to Montgomery form
    uint32_t e[8] = {1u << logn, 0, 0, 0, 0, 0, 0, 0};
    Fr z = fr_sub(fr_pow_limbs(gen, e, 32), fr_one());           // 7^n - 1
    c[0] = g; c[1] = fr_inv(g); c[2] = fr_inv(nn); c[3] = gen; c[4] = fr_inv(gen); c[5] = fr_inv(z);
}
// out[i] = first * base^i for i < count (thread i: square-and-multiply over the bits of i)
__global__ void __launch_bounds__(256) k_pow_table(const Fr *base_p, const Fr *first_p, uint32_t count, Fr *out) {
    __shared__ Fr pw[32];
    if (threadIdx.x == 0) {
        Fr b = *base_p;
        for (int k = 0; k < 32; k++) { pw[k] = b; b = fr_mul(b, b); }
    }
    __syncthreads();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr acc = first_p ? *first_p : fr_one();
    for (int k = 0; k < 32 && (i >> k); k++)
        if ((i >> k) & 1) acc = fr_mul(acc, pw[k]);
    fr_store(&out[i], acc);
}

// out[bitrev(i)] = in[i] * (scale ? scale[i] : 1)
__global__ void __launch_bounds__(256) k_ntt_bitrev(const Fr *__restrict__ in, Fr *__restrict__ out, uint32_t logn,
                                                    const Fr *__restrict__ scale) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << logn)) return;
    Fr v = fr_load(&in[i]);
    if (scale) v = fr_mul(v, fr_load(&scale[i]));
    uint32_t j = logn ? (__brev(i) >> (32 - logn)) : 0;
    fr_store(&out[j], v);
}

// DIT stages s0+1 .. s0+S on a tile: element index = high << (s0+S) | mid << s0 | low, the CTA owns
// all 2^S values of `mid` for one `high` and TL consecutive `low` values.
// Stage t pairs mid with mid | 2^(t-1); twiddle exponent ((mid & (2^(t-1)-1)) << s0 | low) << (k-s0-t).
template <int TL>
__global__ void __launch_bounds__(256) k_ntt_stages(Fr *__restrict__ data, uint32_t logn, uint32_t s0, uint32_t S,
                                                    const Fr *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char dg_ntt_smem[];
    Fr *sm = reinterpret_cast<Fr *>(dg_ntt_smem);              // [mid][l], 2^S x TL
    const uint32_t nmid = 1u << S, low_tiles = (1u << s0) / TL;
    const uint32_t high = blockIdx.x / low_tiles, low0 = (blockIdx.x % low_tiles) * TL;
    const size_t base = ((size_t)high << (s0 + S)) + low0;
    for (uint32_t e = threadIdx.x; e < nmid * TL; e += blockDim.x) {
        uint32_t mid = e / TL, l = e % TL;
        sm[e] = fr_load(&data[base + ((size_t)mid << s0) + l]);
    }
    __syncthreads();
    const uint32_t nbf = (nmid >> 1) * TL;
    for (uint32_t t = 1; t <= S; t++) {
        const uint32_t half = 1u << (t - 1);
        for (uint32_t b = threadIdx.x; b < nbf; b += blockDim.x) {
            uint32_t pair = b / TL, l = b % TL;
            uint32_t jlo = pair & (half - 1);
            uint32_t mid_lo = ((pair >> (t - 1)) << t) | jlo, mid_hi = mid_lo | half;
            uint32_t tw_idx = ((jlo << s0) | (low0 + l)) << (logn - s0 - t);
            Fr u = sm[mid_lo * TL + l];
            Fr v = fr_mul(sm[mid_hi * TL + l], fr_load(&tw[tw_idx]));
            sm[mid_lo * TL + l] = fr_add(u, v);
            sm[mid_hi * TL + l] = fr_sub(u, v);
        }
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < nmid * TL; e += blockDim.x) {
        uint32_t mid = e / TL, l = e % TL;
        fr_store(&data[base + ((size_t)mid << s0) + l], sm[e]);
    }
}

// out[i] = in[i] * (scale ? scale[i] : 1) * (cst ? *cst : 1)
__global__ void __launch_bounds__(256) k_ntt_finish(const Fr *__restrict__ in, Fr *__restrict__ out, uint32_t n,
                                                    const Fr *__restrict__ scale, const Fr *__restrict__ cst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fr_load(&in[i]);
    if (scale) v = fr_mul(v, fr_load(&scale[i]));
    if (cst) v = fr_mul(v, *cst);
    fr_store(&out[i], v);
}
// ab[i] = (a[i] * b[i] - c[i]) * zinv
__global__ void __launch_bounds__(256) k_qap_pointwise(Fr *__restrict__ a, const Fr *__restrict__ b, const Fr *__restrict__ c,
                                                       uint32_t n, const Fr *__restrict__ zinv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_store(&a[i], fr_mul(fr_sub(fr_mul(fr_load(&a[i]), fr_load(&b[i])), fr_load(&c[i])), *zinv));
}

__global__ void __launch_bounds__(128) k_dbg_fr_op(int op, const Fr *a, const Fr *b, uint32_t n, Fr *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = fr_load(&a[i]), y = fr_load(&b[i]), r;
    switch (op) {
        case 0: r = fr_mul(x, y); break;
        case 1: r = fr_add(x, y); break;
        case 2: r = fr_sub(x, y); break;
        default: r = fr_inv(x); break;
    }
    fr_store(&out[i], r);
}

// Fr::into_bigint(): Montgomery -> canonical integer = Montgomery product with the plain integer 1
__global__ void __launch_bounds__(256) k_fr_into_bigint(const Fr *__restrict__ in, Fr *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr one = fr_zero();
    one.l[0] = 1;
    fr_store(&out[i], fr_mul(fr_load(&in[i]), one));
}
int32_t fr_into_bigint_device(const void *in, void *out, size_t n, cudaStream_t s) {
    if (n) DG_LAUNCH(k_fr_into_bigint, div_up(n, 256), 256, 0, s, (const Fr *)in, (Fr *)out, (uint32_t)n);
    return DG_OK;
}

void ntt_release_plans() {
    for (auto &kv : plans()) {
        cudaFree(kv.second.tw_fwd); cudaFree(kv.second.tw_inv); cudaFree(kv.second.scale_fwd); cudaFree(kv.second.scale_inv);
        cudaFree(kv.second.consts);
    }
    plans().clear();
}

static int32_t get_plan(uint32_t logn, cudaStream_t s, NttPlan &out) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = plans().find(logn);
    if (it != plans().end()) { out = it->second; return DG_OK; }
    NttPlan p;
    p.logn = logn;
    size_t n = (size_t)1 << logn, half = n > 1 ? n / 2 : 1;
    Fr *c6 = nullptr;
    DG_CUDA(cudaMalloc(&c6, sizeof(Fr) * 8));
    DG_CUDA(cudaMalloc(&p.tw_fwd, sizeof(Fr) * half));
    DG_CUDA(cudaMalloc(&p.tw_inv, sizeof(Fr) * half));
    DG_CUDA(cudaMalloc(&p.scale_fwd, sizeof(Fr) * n));
    DG_CUDA(cudaMalloc(&p.scale_inv, sizeof(Fr) * n));
    DG_LAUNCH(k_ntt_consts, 1, 32, 0, s, logn, c6);
    DG_LAUNCH(k_pow_table, div_up(half, 256), 256, 0, s, c6 + 0, (const Fr *)nullptr, (uint32_t)half, p.tw_fwd);
    DG_LAUNCH(k_pow_table, div_up(half, 256), 256, 0, s, c6 + 1, (const Fr *)nullptr, (uint32_t)half, p.tw_inv);
    DG_LAUNCH(k_pow_table, div_up(n, 256), 256, 0, s, c6 + 3, (const Fr *)nullptr, (uint32_t)n, p.scale_fwd);
    DG_LAUNCH(k_pow_table, div_up(n, 256), 256, 0, s, c6 + 4, c6 + 2, (uint32_t)n, p.scale_inv);      // n^-1 * 7^-i
    DG_CUDA(cudaStreamSynchronize(s));
    p.consts = c6;                      // [2] = n^-1, [5] = 1/(7^n - 1)
    plans()[logn] = p;
    out = p;
    return DG_OK;
}

// In-place transform of `data` (device), `tmp` is an n-element scratch buffer.
static int32_t ntt_device(Fr *data, Fr *tmp, uint32_t logn, bool inverse, bool coset, cudaStream_t s) {
    if (logn > 28) return fail(DG_ERR_BAD_ARG, "ntt: logn must be <= 28");
    NttPlan p;
    int32_t rc = get_plan(logn, s, p);
    if (rc) return rc;
    const uint32_t n = 1u << logn;
    // bit-reversal (+ forward coset scaling a_i *= 7^i) into tmp, stages in place on tmp, copy / scale back
    DG_LAUNCH(k_ntt_bitrev, div_up(n, 256), 256, 0, s, data, tmp, logn, (!inverse && coset) ? p.scale_fwd : (const Fr *)nullptr);
    const Fr *tw = inverse ? p.tw_inv : p.tw_fwd;
    uint32_t s0 = 0;
    while (s0 < logn) {
        if (s0 == 0) {
            uint32_t S = logn < 8 ? logn : 8;
            uint32_t threads = (1u << S) / 2 < 32 ? 32 : (1u << S) / 2;
            DG_LAUNCH(k_ntt_stages<1>, n >> S, threads, sizeof(Fr) << S, s, tmp, logn, 0u, S, tw);
            s0 = S;
        } else {
            uint32_t S = logn - s0 < 6 ? logn - s0 : 6;
            DG_LAUNCH(k_ntt_stages<8>, n >> (S + 3), 256, (sizeof(Fr) << S) * 8, s, tmp, logn, s0, S, tw);
            s0 += S;
        }
    }
    // copy back, with a_i *= n^-1 (* 7^-i for the coset variant) on the inverse transform
    DG_LAUNCH(k_ntt_finish, div_up(n, 256), 256, 0, s, tmp, data, n, (inverse && coset) ? p.scale_inv : (const Fr *)nullptr,
              (inverse && !coset) ? p.consts + 2 : (const Fr *)nullptr);
    return DG_OK;
}

static int32_t ntt_host(uint8_t *data, uint32_t logn, int32_t inverse, int32_t coset) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!data) return fail(DG_ERR_BAD_ARG, "fr_ntt: null pointer");
    if (logn > 28) return fail(DG_ERR_BAD_ARG, "fr_ntt: logn must be <= 28");
    ThreadState &t = tls();
    size_t n = (size_t)1 << logn;
    rc = t.arena.ensure(2 * Arena::pad(sizeof(Fr) * n), t.stream);
    if (rc) return rc;
    Fr *d = t.arena.alloc<Fr>(n), *tmp = t.arena.alloc<Fr>(n);
    DG_CUDA(cudaMemcpyAsync(d, data, sizeof(Fr) * n, cudaMemcpyHostToDevice, t.stream));
    rc = ntt_device(d, tmp, logn, inverse != 0, coset != 0, t.stream);
    if (rc) return rc;
    DG_CUDA(cudaMemcpyAsync(data, d, sizeof(Fr) * n, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

}  // namespace dg

using namespace dg;

namespace dg {
// Sparse constraint-matrix times assignment: out[i] = sum_k coeff[k] * w[col[k]], k in [row_ptr[i], row_ptr[i+1]).
// The `evaluate_constraint` map at the head of LibsnarkReduction::witness_map_from_matrices
// (legogroth16/src/r1cs_to_qap.rs:150-186) that produces the a, b, c evaluations dg_qap_h_from_abc consumes; CSR
// as crypto_b200/r1cs.py lays the matrices of a Circom circuit out.  One thread per row (R1CS rows are short).
__global__ void __launch_bounds__(128) k_fr_spmv(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                                 const Fr *__restrict__ coeff, uint32_t rows, const Fr *__restrict__ w, uint32_t ncols,
                                                 Fr *__restrict__ out, uint32_t *__restrict__ err_flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    Fr acc = fr_zero();
    for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
        uint32_t c = col[k];
        if (c >= ncols) { atomicOr(err_flag, 2u); continue; }
        acc = fr_add(acc, fr_mul(fr_load(&coeff[k]), fr_load(&w[c])));
    }
    fr_store(&out[i], acc);
}

}  // namespace dg

extern "C" {

int32_t dg_fr_ntt(uint8_t *data, uint32_t logn, int32_t inverse, int32_t coset) { return ntt_host(data, logn, inverse, coset); }

int32_t dg_fr_ntt_device(void *data_dev, void *tmp_dev, uint32_t logn, int32_t inverse, int32_t coset, void *stream) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!data_dev || !tmp_dev) return fail(DG_ERR_BAD_ARG, "fr_ntt_device: null pointer");
    cudaStream_t s = stream ? (cudaStream_t)stream : tls().stream;
    return ntt_device((Fr *)data_dev, (Fr *)tmp_dev, logn, inverse != 0, coset != 0, s);
}

int32_t dg_fr_into_bigint(const uint8_t *fr_mont, size_t n, uint8_t *out_canonical) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (n && (!fr_mont || !out_canonical)) return fail(DG_ERR_BAD_ARG, "fr_into_bigint: null pointer");
    if (n == 0) return DG_OK;
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(32 * n), t.stream);
    if (rc) return rc;
    Fr *d = t.arena.alloc<Fr>(n);
    DG_CUDA(cudaMemcpyAsync(d, fr_mont, 32 * n, cudaMemcpyHostToDevice, t.stream));
    fr_into_bigint_device(d, d, n, t.stream);
    DG_CUDA(cudaMemcpyAsync(out_canonical, d, 32 * n, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

int32_t dg_fr_spmv(const uint32_t *row_ptr, const uint32_t *col, const uint8_t *coeff_mont, size_t rows, size_t nnz, const uint8_t *w_mont,
                   size_t ncols, uint8_t *out_mont) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (rows == 0) return DG_OK;
    if (!row_ptr || !out_mont || (nnz && (!col || !coeff_mont || !w_mont))) return fail(DG_ERR_BAD_ARG, "fr_spmv: null pointer");
    if (rows >= (1ull << 31) || nnz >= (1ull << 32) || ncols >= (1ull << 32)) return fail(DG_ERR_BAD_ARG, "fr_spmv: dimensions too large");
    if (row_ptr[0] != 0 || row_ptr[rows] != nnz) return fail(DG_ERR_BAD_ARG, "fr_spmv: row_ptr must start at 0 and end at nnz");
    for (size_t i = 0; i < rows; i++)
        if (row_ptr[i] > row_ptr[i + 1]) return fail(DG_ERR_BAD_ARG, "fr_spmv: row_ptr must be non-decreasing");
    ThreadState &t = tls();
    rc = t.arena.ensure(Arena::pad(4 * (rows + 1)) + Arena::pad(4 * nnz) + Arena::pad(32 * nnz) + Arena::pad(32 * ncols) + Arena::pad(32 * rows),
                        t.stream);
    if (rc) return rc;
    uint32_t *d_rp = t.arena.alloc<uint32_t>(rows + 1), *d_col = t.arena.alloc<uint32_t>(nnz ? nnz : 1);
    Fr *d_co = t.arena.alloc<Fr>(nnz ? nnz : 1), *d_w = t.arena.alloc<Fr>(ncols ? ncols : 1), *d_o = t.arena.alloc<Fr>(rows);
    DG_CUDA(cudaMemcpyAsync(d_rp, row_ptr, 4 * (rows + 1), cudaMemcpyHostToDevice, t.stream));
    if (nnz) {
        DG_CUDA(cudaMemcpyAsync(d_col, col, 4 * nnz, cudaMemcpyHostToDevice, t.stream));
        DG_CUDA(cudaMemcpyAsync(d_co, coeff_mont, 32 * nnz, cudaMemcpyHostToDevice, t.stream));
        DG_CUDA(cudaMemcpyAsync(d_w, w_mont, 32 * ncols, cudaMemcpyHostToDevice, t.stream));
    }
    DG_LAUNCH(k_fr_spmv, div_up(rows, 128), 128, 0, t.stream, d_rp, d_col, d_co, (uint32_t)rows, d_w, (uint32_t)ncols, d_o, t.err_flag);
    DG_CUDA(cudaMemcpyAsync(out_mont, d_o, 32 * rows, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaMemcpyAsync(t.err_flag_host, t.err_flag, 4, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    if (*t.err_flag_host) {
        cudaMemsetAsync(t.err_flag, 0, 4, t.stream);
        return fail(DG_ERR_BAD_ARG, "fr_spmv: column index out of range");
    }
    return DG_OK;
}

// test hook: out[i] = a[i] (op) b[i]  with op 0 mul, 1 add, 2 sub, 3 inverse(a)
int32_t dg_dbg_fr_op(int32_t op, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!a || !b || !out || n == 0) return fail(DG_ERR_BAD_ARG, "dbg_fr_op: null pointer");
    ThreadState &t = tls();
    rc = t.arena.ensure(3 * Arena::pad(32 * n), t.stream);
    if (rc) return rc;
    Fr *d_a = t.arena.alloc<Fr>(n), *d_b = t.arena.alloc<Fr>(n), *d_o = t.arena.alloc<Fr>(n);
    DG_CUDA(cudaMemcpyAsync(d_a, a, 32 * n, cudaMemcpyHostToDevice, t.stream));
    DG_CUDA(cudaMemcpyAsync(d_b, b, 32 * n, cudaMemcpyHostToDevice, t.stream));
    DG_LAUNCH(k_dbg_fr_op, div_up(n, 128), 128, 0, t.stream, op, d_a, d_b, (uint32_t)n, d_o);
    DG_CUDA(cudaMemcpyAsync(out, d_o, 32 * n, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

// ---- device-chained LegoGroth16 prover (rows a18 + f1) ------------------------------------------------------------------
// The constraint matrices are fixed per circuit, like the proving key: they are uploaded once (dg_r1cs_upload) as one
// device blob of three CSR matrices.  dg_groth16_prove_msms then runs, for one assignment and without leaving the device,
//   witness_map_from_matrices   (legogroth16/src/r1cs_to_qap.rs:150-210: A w, B w, C w, the instance rows of a, 3 iFFT,
//                                3 coset FFT, (ab - c) / Z(7), coset iFFT)
//   into_bigint of h and of the assignment   (prover.rs:281-283, 291-293, 315-317)
//   msm_bigint(h_query, h)      (prover.rs:286)
//   every further MSM of the proof over a contiguous range of the assignment   (prover.rs:299, 326, 334, 344, 363)
// and copies only the 144 / 288-byte results back.
struct R1csMeta { uint64_t off_rp[3], off_col[3], off_co[3], nnz[3], ncons, ninputs, nvars; };

static int32_t lookup_r1cs(uint64_t handle, HandleRec &out) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || it->second.kind != HandleRec::R1CS) return fail(DG_ERR_BAD_ARG, "bad r1cs handle");
    if (it->second.slot != tls().slot) return fail(DG_ERR_BAD_ARG, "the r1cs handle lives on another device");
    out = it->second;
    return DG_OK;
}

int32_t dg_r1cs_upload(const uint32_t *const row_ptr[3], const uint32_t *const col[3], const uint8_t *const coeff_mont[3], size_t num_constraints,
                       size_t num_inputs, size_t num_vars, uint64_t *handle) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!row_ptr || !col || !coeff_mont || !handle) return fail(DG_ERR_BAD_ARG, "r1cs_upload: null pointer");
    if (num_constraints == 0 || num_constraints >= (1ull << 28) || num_vars >= (1ull << 31) || num_inputs == 0 || num_inputs > num_vars)
        return fail(DG_ERR_BAD_ARG, "r1cs_upload: bad dimensions");
    R1csMeta m = {};
    m.ncons = num_constraints; m.ninputs = num_inputs; m.nvars = num_vars;
    size_t total = 0;
    for (int k = 0; k < 3; k++) {
        if (!row_ptr[k]) return fail(DG_ERR_BAD_ARG, "r1cs_upload: null row_ptr");
        if (row_ptr[k][0] != 0) return fail(DG_ERR_BAD_ARG, "r1cs_upload: row_ptr must start at 0");
        for (size_t i = 0; i < num_constraints; i++)
            if (row_ptr[k][i] > row_ptr[k][i + 1]) return fail(DG_ERR_BAD_ARG, "r1cs_upload: row_ptr must be non-decreasing");
        m.nnz[k] = row_ptr[k][num_constraints];
        if (m.nnz[k] && (!col[k] || !coeff_mont[k])) return fail(DG_ERR_BAD_ARG, "r1cs_upload: null col / coeff");
        for (size_t j = 0; j < m.nnz[k]; j++)
            if (col[k][j] >= num_vars) return fail(DG_ERR_BAD_ARG, "r1cs_upload: column index out of range");
        m.off_rp[k] = total; total += Arena::pad(4 * (num_constraints + 1));
        m.off_col[k] = total; total += Arena::pad(4 * (m.nnz[k] + 1));
        m.off_co[k] = total; total += Arena::pad(32 * (m.nnz[k] + 1));
    }
    char *blob = nullptr;
    DG_CUDA(cudaMalloc(&blob, total));
    ThreadState &t = tls();
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 3 && e == cudaSuccess; k++) {
        e = cudaMemcpyAsync(blob + m.off_rp[k], row_ptr[k], 4 * (num_constraints + 1), cudaMemcpyHostToDevice, t.stream);
        if (e == cudaSuccess && m.nnz[k]) e = cudaMemcpyAsync(blob + m.off_col[k], col[k], 4 * m.nnz[k], cudaMemcpyHostToDevice, t.stream);
        if (e == cudaSuccess && m.nnz[k]) e = cudaMemcpyAsync(blob + m.off_co[k], coeff_mont[k], 32 * m.nnz[k], cudaMemcpyHostToDevice, t.stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(t.stream);
    if (e != cudaSuccess) { cudaFree(blob); return fail(DG_ERR_CUDA, cudaGetErrorString(e)); }
    std::lock_guard<std::mutex> lk(ctx().mu);
    uint64_t h = ctx().next_handle++;
    HandleRec r;
    r.kind = HandleRec::R1CS;
    r.dev = blob; r.n = num_constraints; r.slot = t.slot;
    r.meta.assign((const uint64_t *)&m, (const uint64_t *)&m + sizeof(m) / 8);
    ctx().handles[h] = r;
    *handle = h;
    return DG_OK;
}

int32_t dg_r1cs_free(uint64_t handle) {
    void *dev = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx().mu);
        auto it = ctx().handles.find(handle);
        if (it == ctx().handles.end() || it->second.kind != HandleRec::R1CS) return fail(DG_ERR_BAD_ARG, "r1cs_free: bad handle");
        dev = it->second.dev;
        ctx().handles.erase(it);
    }
    cudaDeviceSynchronize();
    cudaFree(dev);
    return DG_OK;
}

int32_t dg_groth16_prove_msms(uint64_t r1cs_handle, const uint8_t *full_assignment_mont, size_t num_vars, uint64_t h_query_handle,
                              const uint64_t *job_bases, const uint64_t *job_offset, const uint64_t *job_count, size_t njobs,
                              uint8_t *out_h_acc_jac, uint8_t *out_jobs_jac, uint8_t *out_h_mont) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!full_assignment_mont || !out_h_acc_jac || (njobs && (!job_bases || !job_offset || !job_count || !out_jobs_jac)))
        return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: null pointer");
    if (njobs > 31) return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: at most 31 MSM jobs per call");
    HandleRec rr;
    rc = lookup_r1cs(r1cs_handle, rr);
    if (rc) return rc;
    R1csMeta m;
    memcpy(&m, rr.meta.data(), sizeof(m));
    if (num_vars != m.nvars) return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: assignment length differs from the circuit's variable count");
    uint32_t logn = 0;
    while (((size_t)1 << logn) < m.ncons + m.ninputs) logn++;                  // Radix2EvaluationDomain::new(num_constraints + num_inputs)
    if (logn > 28) return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: domain too large");
    const size_t D = (size_t)1 << logn;
    // bases of every MSM (copies of the records: the table lock is not held while kernels are queued)
    std::vector<HandleRec> jb(njobs + 1);
    {
        std::lock_guard<std::mutex> lk(ctx().mu);
        for (size_t j = 0; j <= njobs; j++) {
            uint64_t h = j < njobs ? job_bases[j] : h_query_handle;
            auto it = ctx().handles.find(h);
            if (it == ctx().handles.end() || (it->second.kind != HandleRec::BASES_G1 && it->second.kind != HandleRec::BASES_G2))
                return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: bad bases handle");
            if (it->second.slot != tls().slot) return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: a bases handle lives on another device");
            jb[j] = it->second;
        }
    }
    if (jb[njobs].kind != HandleRec::BASES_G1) return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: h_query must be G1");
    const size_t nh = jb[njobs].n < D ? jb[njobs].n : D;                        // msm_bigint truncates to the shorter side
    size_t msm_scratch = 0;
    auto pre_of = [](const HandleRec &r) { return msm_pre_of(r); };
    for (size_t j = 0; j <= njobs; j++) {
        size_t cnt = j < njobs ? job_count[j] : nh;
        if (j < njobs && (job_offset[j] > num_vars || cnt > num_vars - job_offset[j])) return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: scalar range outside the assignment");
        if (cnt > jb[j].n) return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: more scalars than bases");
        size_t b = jb[j].kind == HandleRec::BASES_G2 ? msm_scratch_bytes_g2(cnt, pre_of(jb[j]), false) : msm_scratch_bytes_g1(cnt, pre_of(jb[j]), false);   // the MSMs of a proof overlap each other: no window-group split inside them
        if (b > msm_scratch) msm_scratch = b;
    }
    ThreadState &t = tls();
    cudaStream_t s = t.stream;
    // Several streams (five by default): stream A runs the witness map and then MSMs, streams B and C start on the assignment MSMs at once.
    // The latency-bound tail of one MSM (bucket reduction, window combination: ~1 ms of a 2^18-term MSM) then overlaps the
    // throughput-bound body of another instead of idling the SMs.  (tunable 1 = 2: two streams, the round-2 first form)
    if (!t.stream2) {
        DG_CUDA(cudaStreamCreateWithFlags(&t.stream2, cudaStreamNonBlocking));
        DG_CUDA(cudaEventCreateWithFlags(&t.ev_a, cudaEventDisableTiming));
        DG_CUDA(cudaEventCreateWithFlags(&t.ev_b, cudaEventDisableTiming));
    }
    for (int k = 0; k < 3; k++)
        if (!t.xstream[k]) {
            DG_CUDA(cudaStreamCreateWithFlags(&t.xstream[k], cudaStreamNonBlocking));
            DG_CUDA(cudaEventCreateWithFlags(&t.xev[k], cudaEventDisableTiming));
        }
    // Measured on B200 at D = 2^18 (tools/prover_run.py): 2 streams 12.3 ms, 3 streams 10.7 ms, 5 streams (every MSM of a
    // LegoGroth16 proof on its own stream) 8.6-9.4 ms with resident tables; 15.3 / 13.4 / 10.3 ms with plain keys.
    int NS = ctx().tunable[1].load();                                          // A/B switch: 2 .. 5 streams, default 5
    if (NS < 2 || NS > 5) NS = 5;
    cudaStream_t st[5] = {s, t.stream2, t.xstream[0], t.xstream[1], t.xstream[2]};
    const size_t fixed = 2 * Arena::pad(32 * num_vars) + 4 * Arena::pad(32 * D) + Arena::pad(288 * (njobs + 1));
    {   // one MSM scratch per stream: give up streams rather than fail when a large circuit does not leave room for all of them
        size_t free_b = 0, total_b = 0;
        if (fixed + NS * Arena::pad(msm_scratch) > t.arena.cap && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)   // only when the arena must grow
            while (NS > 2 && fixed + NS * Arena::pad(msm_scratch) > free_b + t.arena.cap) NS--;
    }
    size_t need = fixed + NS * Arena::pad(msm_scratch);
    rc = t.arena.ensure(need, s);
    if (rc) return rc;
    Fr *d_w = t.arena.alloc<Fr>(num_vars), *d_wbig = t.arena.alloc<Fr>(num_vars);
    Fr *d[3], *tmp;
    for (int k = 0; k < 3; k++) d[k] = t.arena.alloc<Fr>(D);
    tmp = t.arena.alloc<Fr>(D);
    uint8_t *d_res = t.arena.alloc<uint8_t>(288 * (njobs + 1));
    char *scratch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < NS; k++) scratch[k] = t.arena.alloc<char>(msm_scratch);
    DG_CUDA(cudaMemcpyAsync(d_w, full_assignment_mont, 32 * num_vars, cudaMemcpyHostToDevice, s));
    fr_into_bigint_device(d_w, d_wbig, num_vars, s);                            // aux / input assignment
    DG_CUDA(cudaEventRecord(t.ev_a, s));
    for (int k = 1; k < NS; k++) DG_CUDA(cudaStreamWaitEvent(st[k], t.ev_a, 0));   // streams B, C may start on the assignment MSMs
    // a, b, c over the domain: constraint rows, then (a only) the instance variables, zero padding
    const char *blob = (const char *)rr.dev;
    for (int k = 0; k < 3; k++) {
        DG_CUDA(cudaMemsetAsync(d[k] + m.ncons, 0, 32 * (D - m.ncons), s));
        DG_LAUNCH(k_fr_spmv, div_up(m.ncons, 128), 128, 0, s, (const uint32_t *)(blob + m.off_rp[k]), (const uint32_t *)(blob + m.off_col[k]),
                  (const Fr *)(blob + m.off_co[k]), (uint32_t)m.ncons, d_w, (uint32_t)num_vars, d[k], t.err_flag + 8);
    }
    DG_CUDA(cudaMemcpyAsync(d[0] + m.ncons, d_w, 32 * m.ninputs, cudaMemcpyDeviceToDevice, s));
    for (int k = 0; k < 3; k++) {
        rc = ntt_device(d[k], tmp, logn, true, false, s);
        if (!rc) rc = ntt_device(d[k], tmp, logn, false, true, s);
        if (rc) return rc;
    }
    NttPlan p;
    rc = get_plan(logn, s, p);
    if (rc) return rc;
    DG_LAUNCH(k_qap_pointwise, div_up(D, 256), 256, 0, s, d[0], d[1], d[2], (uint32_t)D, p.consts + 5);
    rc = ntt_device(d[0], tmp, logn, true, true, s);                            // d[0] = h (Montgomery)
    if (rc) return rc;
    if (out_h_mont) DG_CUDA(cudaMemcpyAsync(out_h_mont, d[0], 32 * D, cudaMemcpyDeviceToHost, s));
    fr_into_bigint_device(d[0], d[1], nh, s);                                   // h_assignment
    // static schedule: the h MSM follows the witness map on stream A; every other MSM goes to the stream with less work
    // queued (cost ~ terms, a G2 term ~2.4 G1 terms; the witness map itself ~ D / 4 terms)
    double load[5] = {(double)nh + (double)D / 4, 0.0, 0.0, 0.0, 0.0};
    int where[32];
    for (size_t j = 0; j < njobs; j++) {
        double cost = (double)job_count[j] * (jb[j].kind == HandleRec::BASES_G2 ? 2.4 : 1.0);
        int k = 0;
        for (int q = 1; q < NS; q++)
            if (load[q] <= load[k]) k = q;
        where[j] = k;
        load[k] += cost;
    }
    where[njobs] = 0;
    uint32_t bad = 0;
    auto run_job = [&](size_t j) -> int32_t {
        const int k = where[j];
        const bool g2 = jb[j].kind == HandleRec::BASES_G2;
        const size_t cnt = j < njobs ? job_count[j] : nh;
        const void *sc = j < njobs ? (const void *)(d_wbig + job_offset[j]) : (const void *)d[1];
        uint32_t *flag = t.err_flag + k;                                     // one flag word per stream
        int32_t r2 = g2 ? msm_run_g2(jb[j].dev, sc, cnt, d_res + 288 * j, scratch[k], flag, st[k], pre_of(jb[j]), nullptr, false)
                        : msm_run_g1(jb[j].dev, sc, cnt, d_res + 288 * j, scratch[k], flag, st[k], pre_of(jb[j]), nullptr, false);
        if (r2) return r2;
        // msm_run clears the flag when it starts: collect it per MSM
        DG_CUDA(cudaMemcpyAsync(t.err_flag_host + 1 + j, flag, 4, cudaMemcpyDeviceToHost, st[k]));
        return DG_OK;
    };
    // the jobs of streams B, C first (they can start now), then stream A's: h, then the rest
    for (size_t j = 0; j < njobs; j++)
        if (where[j] != 0 && (rc = run_job(j))) return rc;
    if ((rc = run_job(njobs))) return rc;
    for (size_t j = 0; j < njobs; j++)
        if (where[j] == 0 && (rc = run_job(j))) return rc;
    DG_CUDA(cudaEventRecord(t.ev_b, st[1]));
    DG_CUDA(cudaStreamWaitEvent(s, t.ev_b, 0));
    for (int k = 2; k < NS; k++) {
        DG_CUDA(cudaEventRecord(t.xev[k - 2], st[k]));
        DG_CUDA(cudaStreamWaitEvent(s, t.xev[k - 2], 0));
    }
    for (size_t j = 0; j < njobs; j++)
        DG_CUDA(cudaMemcpyAsync(out_jobs_jac + 288 * j, d_res + 288 * j, jb[j].kind == HandleRec::BASES_G2 ? 288 : 144, cudaMemcpyDeviceToHost, s));
    DG_CUDA(cudaMemcpyAsync(out_h_acc_jac, d_res + 288 * njobs, 144, cudaMemcpyDeviceToHost, s));
    DG_CUDA(cudaStreamSynchronize(s));
    for (size_t j = 0; j <= njobs && j < 32; j++) bad |= t.err_flag_host[1 + j];
    if (bad) return fail(DG_ERR_BAD_ARG, "groth16_prove_msms: assignment element is not a reduced Montgomery residue");
    return DG_OK;
}

// h = coset_ifft( (coset_fft(ifft a) * coset_fft(ifft b) - coset_fft(ifft c)) / (7^n - 1) )
int32_t dg_qap_h_from_abc(const uint8_t *a, const uint8_t *b, const uint8_t *c, uint32_t logn, uint8_t *out_h) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!a || !b || !c || !out_h) return fail(DG_ERR_BAD_ARG, "qap_h_from_abc: null pointer");
    if (logn > 28) return fail(DG_ERR_BAD_ARG, "qap_h_from_abc: logn must be <= 28");
    ThreadState &t = tls();
    size_t n = (size_t)1 << logn;
    rc = t.arena.ensure(4 * Arena::pad(sizeof(Fr) * n), t.stream);
    if (rc) return rc;
    Fr *d[3], *tmp;
    for (int k = 0; k < 3; k++) d[k] = t.arena.alloc<Fr>(n);
    tmp = t.arena.alloc<Fr>(n);
    const uint8_t *src[3] = {a, b, c};
    for (int k = 0; k < 3; k++) DG_CUDA(cudaMemcpyAsync(d[k], src[k], sizeof(Fr) * n, cudaMemcpyHostToDevice, t.stream));
    for (int k = 0; k < 3; k++) {
        rc = ntt_device(d[k], tmp, logn, true, false, t.stream);
        if (rc) return rc;
        rc = ntt_device(d[k], tmp, logn, false, true, t.stream);
        if (rc) return rc;
    }
    NttPlan p;
    rc = get_plan(logn, t.stream, p);
    if (rc) return rc;
    DG_LAUNCH(k_qap_pointwise, div_up(n, 256), 256, 0, t.stream, d[0], d[1], d[2], (uint32_t)n, p.consts + 5);
    rc = ntt_device(d[0], tmp, logn, true, true, t.stream);
    if (rc) return rc;
    DG_CUDA(cudaMemcpyAsync(out_h, d[0], sizeof(Fr) * n, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    return DG_OK;
}

}  // extern "C"
