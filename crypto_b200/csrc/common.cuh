// Host-side plumbing shared by the translation units of libdockgpu.so: error reporting,
// per-thread stream + scratch arena (the C ABI is re-entrant from rayon-style worker threads,
// SURVEY.md 8b "Threading"), handle table, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include "../../include/dockgpu.h"

namespace dg {

#define DG_MAX_DEVICES 16

struct HandleRec {
    enum Kind { BASES_G1, BASES_G2, TABLE_G1, TABLE_G2, SHARDED_G1, SHARDED_G2, R1CS } kind;
    void *dev = nullptr;
    size_t n = 0;            // bases: point count; tables: total records
    int window = 0, nwin = 0;   // tables: window geometry; bases: precompute window / rows (0 = plain)
    int slot = 0;            // index into Context::devices of the GPU that owns `dev`
    size_t phi_off = 0;      // plain bases: dev holds 2 n records, phi(P_i) = (beta x_i, y_i) at phi_off + i (GLV, msm_kernels.cuh)
    // SHARDED_*: one contiguous base range per device of dg_init_devices; shard d covers
    // [shard_lo[d], shard_lo[d + 1]) and lives behind the ordinary single-device handle shard_handle[d]
    std::vector<uint64_t> shard_handle;
    std::vector<size_t> shard_lo;
    std::vector<uint64_t> meta;      // R1CS: layout of the device blob (ntt.cu R1csMeta)
};

// One host thread per device of dg_init_devices: it owns that device's stream and scratch arena (its
// thread-local ThreadState), so a sharded call is N ordinary single-device calls issued concurrently.
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    bool stop = false;
    void *partial = nullptr;          // 288 B device buffer: this device's partial MSM result (read by the fold over peer memory)
};

struct Context {
    std::mutex mu;
    bool inited = false;
    int device = 0;                               // devices[0]: the device every single-GPU entry point runs on
    int ndev = 0;
    int devices[DG_MAX_DEVICES] = {};
    bool peer[DG_MAX_DEVICES] = {};               // devices[0] can load from devices[d] directly (NVLink peer access)
    Worker *workers[DG_MAX_DEVICES] = {};
    int sm_count = 148;
    std::mutex sharded_mu;                        // one sharded MSM at a time (the per-device partial buffers are shared)
    std::mutex attr_mu;
    std::set<std::pair<int, const void *>> attr_done;   // (device, kernel) pairs whose dynamic shared memory limit is raised
    std::unordered_map<uint64_t, HandleRec> handles;
    uint64_t next_handle = 1;
    std::atomic<uint64_t> launches{0};
    std::atomic<int> msm_window_override{0};
    std::atomic<int> msm_rounds_override{-1};   // batch-affine rounds: -1 = automatic
    std::atomic<int> tunable[8] = {};             // dg_dbg_set_tunable: 0 = minimum waves per batch-affine round, 2 = window-group split, 3 = max outputs per thread (0 = defaults)
    // optional per-kernel timing of the dominant kernel (bench.py roofline): event pairs recorded
    // on the launching stream around every k_accumulate launch while enabled
    std::atomic<int> prof_enabled{0};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
};
Context &ctx();

// Grow-only device scratch; contents are zeroed before release because scalars may be secret
// material (vb_accumulator/src/positive.rs:349-352 zeroizes them on the CPU side).
// Asynchronous users (the *_device entry points return without synchronising) call mark(): the next ensure() on a
// different stream then waits for that event before handing the same memory out, and the grow path waits for it on
// the host before freeing.
struct Arena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    cudaStream_t last_stream = nullptr;
    cudaEvent_t last_event = nullptr;
    bool pending = false;
    int32_t ensure(size_t bytes, cudaStream_t s);
    int32_t mark(cudaStream_t s);
    void reset() { used = 0; }
    template <class T> T *alloc(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        char *p = base + used;
        used += bytes;
        return reinterpret_cast<T *>(p);
    }
    static size_t pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
    void release();
};

struct ThreadState {
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // second stream for calls that overlap independent MSMs (dg_groth16_prove_msms); created on first use
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    cudaStream_t xstream[3] = {nullptr, nullptr, nullptr};   // further streams of the chained prover (streams C, D, E)
    cudaEvent_t xev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[4] = {};     // chunked scalar staging of the host MSM path (capi.cu msm_host)
    cudaStream_t split_stream = nullptr;   // high-priority stream of the window-group split (msm_host.cuh msm_run); created on first use
    cudaEvent_t split_ev[2] = {nullptr, nullptr};
    int slot = 0;                     // device slot this thread drives: 0 for callers, d for the worker of devices[d]
    Arena arena;
    std::string err;
    uint32_t *err_flag = nullptr;     // device word the kernels OR error bits into
    uint32_t *err_flag_host = nullptr;  // pinned mirror
    ~ThreadState();
};
ThreadState &tls();

int32_t fail(int32_t code, const std::string &msg);
int32_t check_init();
// Raises a kernel's dynamic shared memory limit once per (device, kernel); safe from concurrent callers.
int32_t func_smem_opt_in(const void *func, size_t bytes);
template <class K> static inline int32_t smem_opt_in(K kernel, size_t bytes) { return func_smem_opt_in((const void *)kernel, bytes); }
// Runs fn(slot) on the worker thread of every device slot in [0, nslots) concurrently; returns the first failure
// (its message becomes the caller's dg_last_error).
int32_t run_on_devices(int nslots, const std::function<int32_t(int)> &fn);

#define DG_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return dg::fail(e_ == cudaErrorMemoryAllocation ? DG_ERR_OOM : DG_ERR_CUDA,            \
                            std::string(#call) + ": " + cudaGetErrorString(e_));                   \
    } while (0)

#define DG_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
    do {                                                                                           \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                \
        dg::ctx().launches.fetch_add(1, std::memory_order_relaxed);                                \
    } while (0)

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- internal entry points implemented per translation unit (device pointers, async) ----------
// Precomputed-bases descriptor: c = 0 means plain bases; otherwise the table holds ceil(256/c)
// rows of row_stride affine points, row k = 2^(c*k) * P_i (dg_bases_precompute).
// phi_off != 0 (plain bases only): the array already holds the GLV images phi(P_i) at index phi_off + i (resident handles
// are expanded once at upload); 0 = msm_run expands the bases into its scratch first.
struct MsmPre { int c; uint32_t row_stride; uint32_t phi_off; };
static inline MsmPre msm_pre_of(const HandleRec &r) {
    return r.window ? MsmPre{r.window, (uint32_t)r.n, 0u} : MsmPre{0, 0u, (uint32_t)r.phi_off};
}
// digits per GLV half-scalar (k < 2^127): the top digit stays below 2^(c-1) even with the carry
static inline int glv_ndigits(int c) { return (128 + c - 1) / c; }
int32_t glv_expand_g1(const void *in, size_t n, void *out, size_t phi_off, cudaStream_t s);      // msm_g1.cu
int32_t glv_expand_g2(const void *in, size_t n, void *out, size_t phi_off, cudaStream_t s);      // msm_g2.cu
// Signed radix-2^c digits per scalar.  k_digits first maps a canonical scalar s to min(s, r - s) < 2^254 (negating
// the point), so ceil(254 / c) digits suffice: the top digit is narrower than c bits and stays <= 2^(c-1) even with
// the carry -- except when c divides 254, where it is full width and its carry needs one more digit.
static inline int msm_ndigits(int c) { return (254 + c - 1) / c + (254 % c == 0 ? 1 : 0); }
// allow_split = false: the run stays on the caller's stream alone (no window-group split, msm_host.cuh); the scratch size
// and the run must be asked with the same value
size_t msm_scratch_bytes_g1(size_t n, MsmPre pre, bool allow_split = true);
size_t msm_scratch_bytes_g2(size_t n, MsmPre pre, bool allow_split = true);
void msm_plan_g1(size_t n, MsmPre pre, int *c, int *rounds);
void msm_plan_g2(size_t n, MsmPre pre, int *c, int *rounds);
// Scalars that are still arriving from the host: chunk k = [lo[k], lo[k + 1]) is on the device once ev[k] has fired (the
// copies run on a second stream).  The histogram pass of the digit kernel is launched per chunk behind its event, so the
// PCIe transfer of chunk k + 1 overlaps the counting of chunk k.
struct MsmStage { int nchunks; cudaEvent_t ev[4]; size_t lo[5]; };
int32_t msm_run_g1(const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, char *scratch,
                   uint32_t *err_flag, cudaStream_t s, MsmPre pre, const MsmStage *stage = nullptr, bool allow_split = true);
int32_t msm_run_g2(const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, char *scratch,
                   uint32_t *err_flag, cudaStream_t s, MsmPre pre, const MsmStage *stage = nullptr, bool allow_split = true);
void ntt_release_plans();                                                                  // ntt.cu
int32_t fr_into_bigint_device(const void *in, void *out, size_t n, cudaStream_t s);   // ntt.cu
// Partial results of a sharded MSM, one Jacobian record per device; entries may point into peer memory (NVLink).
struct PtrList { const void *p[DG_MAX_DEVICES]; };
int32_t fold_ptrs_g1(const PtrList &pl, int k, void *out_jac_dev, cudaStream_t s);       // batch_g1.cu
int32_t fold_ptrs_g2(const PtrList &pl, int k, void *out_jac_dev, cudaStream_t s);       // batch_g2.cu
int32_t bases_precompute_g1(HandleRec &rec, int c, cudaStream_t s);
int32_t bases_precompute_g2(HandleRec &rec, int c, cudaStream_t s);

}  // namespace dg
