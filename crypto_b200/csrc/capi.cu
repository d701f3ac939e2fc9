// C ABI of libdockgpu.so (include/dockgpu.h): context, handles, host<->device staging and the
// MSM entry points.  Everything here is plumbing around the kernels; no arithmetic on the host.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "ec.cuh"

namespace dg {

Context &ctx() {
    static Context c;
    return c;
}
ThreadState &tls() {
    static thread_local ThreadState t;
    return t;
}
ThreadState::~ThreadState() {
    // Process teardown order vs. the CUDA runtime is undefined: only release when still usable.
    if (stream && cudaStreamQuery(stream) != cudaErrorCudartUnloading) {
        arena.release();
        if (err_flag) cudaFree(err_flag);
        if (err_flag_host) cudaFreeHost(err_flag_host);
        cudaStreamDestroy(stream);
    }
}
int32_t fail(int32_t code, const std::string &msg) {
    tls().err = msg;
    return code;
}
int32_t check_init() {
    Context &c = ctx();
    if (!c.inited) return fail(DG_ERR_NOT_INIT, "dg_init has not been called");
    ThreadState &t = tls();
    if (!t.stream) {
        DG_CUDA(cudaSetDevice(c.device));
        DG_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
        DG_CUDA(cudaMalloc(&t.err_flag, 256));
        DG_CUDA(cudaMemset(t.err_flag, 0, 256));
        DG_CUDA(cudaMallocHost(&t.err_flag_host, 256));
    }
    return DG_OK;
}
int32_t Arena::ensure(size_t bytes, cudaStream_t s) {
    used = 0;
    if (bytes <= cap) return DG_OK;
    DG_CUDA(cudaStreamSynchronize(s));
    release();
    size_t want = bytes + bytes / 8 + (1 << 20);
    DG_CUDA(cudaMalloc(&base, want));
    cap = want;
    return DG_OK;
}
void Arena::release() {
    if (base) {
        cudaMemset(base, 0, cap);       // scalars may be secret material
        cudaFree(base);
    }
    base = nullptr;
    cap = 0;
}

static int32_t read_err_flag(ThreadState &t, const char *what) {
    DG_CUDA(cudaMemcpyAsync(t.err_flag_host, t.err_flag, 4, cudaMemcpyDeviceToHost, t.stream));
    DG_CUDA(cudaStreamSynchronize(t.stream));
    if (*t.err_flag_host) {
        cudaMemsetAsync(t.err_flag, 0, 4, t.stream);
        return fail(DG_ERR_BAD_ARG, std::string(what) + ": scalar is not canonical (>= r)");
    }
    return DG_OK;
}

template <bool G2>
static int32_t msm_host(uint64_t handle, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac,
                        bool mont_scalars = false) {
    int32_t rc = check_init();
    if (rc) return rc;
    const size_t PT = G2 ? 192 : 96, JAC = G2 ? 288 : 144;
    if (!out_jac || (n && !scalars)) return fail(DG_ERR_BAD_ARG, "msm: null pointer");
    if ((handle != 0) == (bases != nullptr) && n) return fail(DG_ERR_BAD_ARG, "msm: give exactly one of bases_handle / bases");
    ThreadState &t = tls();
    const void *bases_dev = nullptr;
    MsmPre pre = {0, 0};
    if (handle) {
        std::lock_guard<std::mutex> lk(ctx().mu);
        auto it = ctx().handles.find(handle);
        if (it == ctx().handles.end() || it->second.kind != (G2 ? HandleRec::BASES_G2 : HandleRec::BASES_G1))
            return fail(DG_ERR_BAD_ARG, "msm: bad bases handle");
        if (n > it->second.n) return fail(DG_ERR_BAD_ARG, "msm: n exceeds uploaded bases");
        bases_dev = it->second.dev;
        if (it->second.window) pre = {it->second.window, (uint32_t)it->second.n};
    }
    size_t need = (G2 ? msm_scratch_bytes_g2(n, pre) : msm_scratch_bytes_g1(n, pre)) + Arena::pad(32 * n) + Arena::pad(JAC) +
                  (handle ? 0 : Arena::pad(PT * n));
    rc = t.arena.ensure(need, t.stream);
    if (rc) return rc;
    uint8_t *d_scalars = t.arena.alloc<uint8_t>(32 * n);
    uint8_t *d_out = t.arena.alloc<uint8_t>(JAC);
    if (n) DG_CUDA(cudaMemcpyAsync(d_scalars, scalars, 32 * n, cudaMemcpyHostToDevice, t.stream));
    if (n && mont_scalars) fr_into_bigint_device(d_scalars, d_scalars, n, t.stream);   // msm_unchecked: into_bigint first
    if (!handle && n) {
        uint8_t *d_bases = t.arena.alloc<uint8_t>(PT * n);
        DG_CUDA(cudaMemcpyAsync(d_bases, bases, PT * n, cudaMemcpyHostToDevice, t.stream));
        bases_dev = d_bases;
    }
    char *scratch = t.arena.alloc<char>(need - t.arena.used);
    rc = G2 ? msm_run_g2(bases_dev, d_scalars, n, d_out, scratch, t.err_flag, t.stream, pre)
            : msm_run_g1(bases_dev, d_scalars, n, d_out, scratch, t.err_flag, t.stream, pre);
    if (rc) return rc;
    DG_CUDA(cudaMemcpyAsync(out_jac, d_out, JAC, cudaMemcpyDeviceToHost, t.stream));
    return read_err_flag(t, "msm");
}

template <bool G2>
static int32_t msm_device(uint64_t handle, const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, void *stream) {
    int32_t rc = check_init();
    if (rc) return rc;
    MsmPre pre = {0, 0};
    if (handle) {
        std::lock_guard<std::mutex> lk(ctx().mu);
        auto it = ctx().handles.find(handle);
        if (it == ctx().handles.end() || it->second.kind != (G2 ? HandleRec::BASES_G2 : HandleRec::BASES_G1))
            return fail(DG_ERR_BAD_ARG, "msm_device: bad bases handle");
        if (n > it->second.n) return fail(DG_ERR_BAD_ARG, "msm_device: n exceeds uploaded bases");
        bases_dev = it->second.dev;
        if (it->second.window) pre = {it->second.window, (uint32_t)it->second.n};
    }
    if (!out_jac_dev || (n && (!bases_dev || !scalars_dev))) return fail(DG_ERR_BAD_ARG, "msm_device: null pointer");
    ThreadState &t = tls();
    cudaStream_t s = stream ? (cudaStream_t)stream : t.stream;
    size_t need = G2 ? msm_scratch_bytes_g2(n, pre) : msm_scratch_bytes_g1(n, pre);
    rc = t.arena.ensure(need, s);
    if (rc) return rc;
    char *scratch = t.arena.alloc<char>(need);
    return G2 ? msm_run_g2(bases_dev, scalars_dev, n, out_jac_dev, scratch, t.err_flag, s, pre)
              : msm_run_g1(bases_dev, scalars_dev, n, out_jac_dev, scratch, t.err_flag, s, pre);
}

template <bool G2> static int32_t bases_upload(const uint8_t *affine, size_t n, uint64_t *handle) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!affine || !handle || n == 0) return fail(DG_ERR_BAD_ARG, "bases_upload: null pointer or n == 0");
    const size_t PT = G2 ? 192 : 96;
    void *dev = nullptr;
    DG_CUDA(cudaMalloc(&dev, PT * n));
    ThreadState &t = tls();
    cudaError_t e = cudaMemcpyAsync(dev, affine, PT * n, cudaMemcpyHostToDevice, t.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(t.stream);
    if (e != cudaSuccess) { cudaFree(dev); return fail(DG_ERR_CUDA, cudaGetErrorString(e)); }
    std::lock_guard<std::mutex> lk(ctx().mu);
    uint64_t h = ctx().next_handle++;
    HandleRec r;
    r.kind = G2 ? HandleRec::BASES_G2 : HandleRec::BASES_G1;
    r.dev = dev; r.n = n;
    ctx().handles[h] = r;
    *handle = h;
    return DG_OK;
}

}  // namespace dg

using namespace dg;

extern "C" {

int32_t dg_init(int32_t device) {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.inited) return DG_OK;
    int count = 0;
    DG_CUDA(cudaGetDeviceCount(&count));
    if (count == 0) return fail(DG_ERR_CUDA, "no CUDA device visible: libdockgpu has no CPU fallback");
    if (device < 0) DG_CUDA(cudaGetDevice(&device));
    if (device >= count) return fail(DG_ERR_BAD_ARG, "dg_init: device index out of range");
    DG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(DG_ERR_CUDA, "libdockgpu is built for sm_100a (B200) only");
    // The hot kernels gather 96-byte records at random from tables far larger than L2; ask for the
    // smallest DRAM->L2 fetch granularity so a miss does not drag the whole 128-byte line in.
    {
        size_t gran = 32;
        if (const char *e = getenv("DG_L2_FETCH_GRANULARITY")) gran = (size_t)atoi(e);
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    c.inited = true;
    return DG_OK;
}

int32_t dg_shutdown(void) {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.inited) return DG_OK;
    cudaDeviceSynchronize();
    for (auto &kv : c.handles) cudaFree(kv.second.dev);
    c.handles.clear();
    ntt_release_plans();
    ThreadState &t = tls();
    t.arena.release();
    c.inited = false;
    return DG_OK;
}

int32_t dg_last_error(char *buf, size_t cap) {
    if (!buf || cap == 0) return DG_ERR_BAD_ARG;
    const std::string &e = tls().err;
    size_t k = e.size() < cap - 1 ? e.size() : cap - 1;
    memcpy(buf, e.data(), k);
    buf[k] = 0;
    return DG_OK;
}

uint64_t dg_launch_count(void) { return ctx().launches.load(); }

int32_t dg_sync(void) {
    int32_t rc = check_init();
    if (rc) return rc;
    DG_CUDA(cudaStreamSynchronize(tls().stream));
    return DG_OK;
}

int32_t dg_bases_upload_g1(const uint8_t *affine, size_t n, uint64_t *handle) { return bases_upload<false>(affine, n, handle); }
int32_t dg_bases_upload_g2(const uint8_t *affine, size_t n, uint64_t *handle) { return bases_upload<true>(affine, n, handle); }
int32_t dg_bases_free(uint64_t handle) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || (it->second.kind != HandleRec::BASES_G1 && it->second.kind != HandleRec::BASES_G2))
        return fail(DG_ERR_BAD_ARG, "bases_free: bad handle");
    cudaDeviceSynchronize();
    cudaFree(it->second.dev);
    ctx().handles.erase(it);
    return DG_OK;
}

int32_t dg_msm_g1(uint64_t h, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out) { return msm_host<false>(h, bases, scalars, n, out); }
int32_t dg_msm_g2(uint64_t h, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out) { return msm_host<true>(h, bases, scalars, n, out); }
int32_t dg_msm_unchecked_g1(uint64_t h, const uint8_t *bases, const uint8_t *fr_mont, size_t n, uint8_t *out) { return msm_host<false>(h, bases, fr_mont, n, out, true); }
int32_t dg_msm_unchecked_g2(uint64_t h, const uint8_t *bases, const uint8_t *fr_mont, size_t n, uint8_t *out) { return msm_host<true>(h, bases, fr_mont, n, out, true); }
int32_t dg_msm_g1_device(const void *b, const void *s, size_t n, void *o, void *st) { return msm_device<false>(0, b, s, n, o, st); }
int32_t dg_msm_g2_device(const void *b, const void *s, size_t n, void *o, void *st) { return msm_device<true>(0, b, s, n, o, st); }
int32_t dg_msm_g1_handle_device(uint64_t h, const void *s, size_t n, void *o, void *st) {
    if (!h) return fail(DG_ERR_BAD_ARG, "msm_handle_device: null handle");
    return msm_device<false>(h, nullptr, s, n, o, st);
}
int32_t dg_msm_g2_handle_device(uint64_t h, const void *s, size_t n, void *o, void *st) {
    if (!h) return fail(DG_ERR_BAD_ARG, "msm_handle_device: null handle");
    return msm_device<true>(h, nullptr, s, n, o, st);
}
int32_t dg_bases_precompute(uint64_t handle, int32_t c) {
    int32_t rc = check_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || (it->second.kind != HandleRec::BASES_G1 && it->second.kind != HandleRec::BASES_G2))
        return fail(DG_ERR_BAD_ARG, "bases_precompute: bad handle");
    if (c == 0) c = it->second.n >= (1u << 22) ? 20 : 17;     // 13 / 15 rows; measured with the batch-affine stage (tools/sweep_rounds.py)
    DG_CUDA(cudaDeviceSynchronize());
    return it->second.kind == HandleRec::BASES_G2 ? bases_precompute_g2(it->second, c, tls().stream)
                                                  : bases_precompute_g1(it->second, c, tls().stream);
}
int32_t dg_prof_enable(int32_t on) {
    ctx().prof_enabled.store(on ? 1 : 0);
    return DG_OK;
}
int32_t dg_prof_read_accumulate(double *mean_ms, int32_t *count) {
    if (!mean_ms || !count) return fail(DG_ERR_BAD_ARG, "prof_read: null pointer");
    DG_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(ctx().mu);
    double tot = 0;
    int n = 0;
    for (auto &pr : ctx().prof_events) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { tot += ms; n++; }
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    ctx().prof_events.clear();
    *mean_ms = n ? tot / n : 0.0;
    *count = n;
    return DG_OK;
}
int32_t dg_msm_set_window(int32_t c) {
    if (c < 0 || c == 1 || c > 24) return fail(DG_ERR_BAD_ARG, "msm_set_window: c must be 0 or in [2, 24]");
    ctx().msm_window_override.store(c);
    return DG_OK;
}

int32_t dg_msm_set_affine_rounds(int32_t rounds) {
    if (rounds < -1 || rounds > 10) return fail(DG_ERR_BAD_ARG, "msm_set_affine_rounds: rounds must be -1 (automatic) or in [0, 10]");
    ctx().msm_rounds_override.store(rounds);
    return DG_OK;
}
int32_t dg_dbg_set_tunable(int32_t id, int32_t value) {
    if (id < 0 || id >= 8) return fail(DG_ERR_BAD_ARG, "dbg_set_tunable: id out of range");
    ctx().tunable[id].store(value);
    return DG_OK;
}
int32_t dg_msm_plan(size_t n, int32_t is_g2, int32_t precomputed_window_bits, int32_t *window_bits, int32_t *affine_rounds) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!window_bits || !affine_rounds) return fail(DG_ERR_BAD_ARG, "msm_plan: null pointer");
    MsmPre pre = {precomputed_window_bits, precomputed_window_bits ? (uint32_t)n : 0u};
    int c = 0, r = 0;
    if (is_g2) msm_plan_g2(n, pre, &c, &r); else msm_plan_g1(n, pre, &c, &r);
    *window_bits = c;
    *affine_rounds = r;
    return DG_OK;
}

}  // extern "C"
