// C ABI of libdockgpu.so (include/dockgpu.h): context, handles, host<->device staging and the
// MSM entry points.  Everything here is plumbing around the kernels; no arithmetic on the host.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "ec.cuh"

namespace dg {

Context &ctx() {
    static Context *c = new Context;      // never destroyed: worker threads and late callers may outlive static destruction
    return *c;
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
        if (stream2) cudaStreamDestroy(stream2);
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
        for (int k = 0; k < 3; k++) {
            if (xstream[k]) cudaStreamDestroy(xstream[k]);
            if (xev[k]) cudaEventDestroy(xev[k]);
        }
        for (int k = 0; k < 4; k++)
            if (stage_ev[k]) cudaEventDestroy(stage_ev[k]);
        if (split_stream) cudaStreamDestroy(split_stream);
        for (int k = 0; k < 2; k++)
            if (split_ev[k]) cudaEventDestroy(split_ev[k]);
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
        DG_CUDA(cudaSetDevice(c.devices[t.slot]));
        DG_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
        DG_CUDA(cudaMalloc(&t.err_flag, 256));
        DG_CUDA(cudaMemset(t.err_flag, 0, 256));
        DG_CUDA(cudaMallocHost(&t.err_flag_host, 256));
    }
    return DG_OK;
}
int32_t func_smem_opt_in(const void *func, size_t bytes) {
    Context &c = ctx();
    int dev = 0;
    DG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(c.attr_mu);
    auto key = std::make_pair(dev, func);
    if (c.attr_done.count(key)) return DG_OK;
    DG_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    c.attr_done.insert(key);
    return DG_OK;
}

int32_t Arena::ensure(size_t bytes, cudaStream_t s) {
    used = 0;
    if (pending && last_stream != s && bytes <= cap) DG_CUDA(cudaStreamWaitEvent(s, last_event, 0));   // another stream may still be using this memory
    if (bytes <= cap) return DG_OK;
    if (pending) DG_CUDA(cudaEventSynchronize(last_event));
    DG_CUDA(cudaStreamSynchronize(s));
    pending = false;
    release();
    size_t want = bytes + bytes / 8 + (1 << 20);
    DG_CUDA(cudaMalloc(&base, want));
    cap = want;
    return DG_OK;
}
int32_t Arena::mark(cudaStream_t s) {
    if (!last_event) DG_CUDA(cudaEventCreateWithFlags(&last_event, cudaEventDisableTiming));
    DG_CUDA(cudaEventRecord(last_event, s));
    last_stream = s;
    pending = true;
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

// ---- per-device worker threads (single-process multi-GPU) ---------------------------
static void worker_main(Worker *w, int slot) {
    tls().slot = slot;
    cudaSetDevice(ctx().devices[slot]);
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->stop || !w->q.empty(); });
            if (w->q.empty()) return;                      // stop requested and nothing left to run
            job = std::move(w->q.front());
            w->q.pop_front();
        }
        job();
    }
}
static void ensure_workers() {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    for (int d = 0; d < c.ndev; d++)
        if (!c.workers[d]) {
            Worker *w = new Worker;
            w->th = std::thread(worker_main, w, d);
            c.workers[d] = w;
        }
}
static void stop_workers() {                               // caller holds ctx().mu
    Context &c = ctx();
    for (int d = 0; d < DG_MAX_DEVICES; d++) {
        Worker *w = c.workers[d];
        if (!w) continue;
        {
            std::lock_guard<std::mutex> lk(w->mu);
            if (w->partial) {
                void *p = w->partial;
                w->q.emplace_back([p] { cudaFree(p); });
            }
            w->stop = true;
        }
        w->cv.notify_all();
        w->th.join();
        delete w;
        c.workers[d] = nullptr;
    }
}
int32_t run_on_devices(int nslots, const std::function<int32_t(int)> &fn) {
    Context &c = ctx();
    if (nslots > c.ndev) return fail(DG_ERR_BAD_ARG, "run_on_devices: more slots than devices");
    ensure_workers();
    struct Join {
        std::mutex mu;
        std::condition_variable cv;
        int left;
        int32_t rc[DG_MAX_DEVICES];
        std::string msg[DG_MAX_DEVICES];
    } j;
    j.left = nslots;
    for (int d = 0; d < nslots; d++) {
        Worker *w = c.workers[d];
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->q.emplace_back([&j, &fn, d] {
                int32_t rc = check_init();
                if (!rc) rc = fn(d);
                std::string m = rc ? tls().err : std::string();
                std::lock_guard<std::mutex> lk2(j.mu);
                j.rc[d] = rc;
                j.msg[d] = m;
                if (--j.left == 0) j.cv.notify_one();
            });
        }
        w->cv.notify_one();
    }
    {
        std::unique_lock<std::mutex> lk(j.mu);
        j.cv.wait(lk, [&] { return j.left == 0; });
    }
    for (int d = 0; d < nslots; d++)
        if (j.rc[d]) return fail(j.rc[d], "device " + std::to_string(c.devices[d]) + ": " + j.msg[d]);
    return DG_OK;
}

static int32_t read_err_flag(ThreadState &t, cudaStream_t s, const char *what) {
    DG_CUDA(cudaMemcpyAsync(t.err_flag_host, t.err_flag, 4, cudaMemcpyDeviceToHost, s));
    DG_CUDA(cudaStreamSynchronize(s));
    if (*t.err_flag_host) {
        cudaMemsetAsync(t.err_flag, 0, 4, s);
        return fail(DG_ERR_BAD_ARG, std::string(what) + ": scalar is not canonical (>= r)");
    }
    return DG_OK;
}

// Looks a bases handle up for the calling thread's device.  Copies the record: the table lock is not held
// while kernels are queued.
template <bool G2> static int32_t lookup_bases(uint64_t handle, size_t n, const char *what, HandleRec &out) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || it->second.kind != (G2 ? HandleRec::BASES_G2 : HandleRec::BASES_G1))
        return fail(DG_ERR_BAD_ARG, std::string(what) + ": bad bases handle");
    if (it->second.slot != tls().slot) return fail(DG_ERR_BAD_ARG, std::string(what) + ": the handle lives on another device");
    if (n > it->second.n) return fail(DG_ERR_BAD_ARG, std::string(what) + ": n exceeds uploaded bases");
    out = it->second;
    return DG_OK;
}

// out_keep_dev != nullptr: the Jacobian result stays in that device buffer and nothing is copied to the host
// (the per-device leg of a sharded MSM); otherwise it is copied to out_jac.
template <bool G2>
static int32_t msm_host(uint64_t handle, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac,
                        bool mont_scalars = false, void *out_keep_dev = nullptr) {
    int32_t rc = check_init();
    if (rc) return rc;
    const size_t PT = G2 ? 192 : 96, JAC = G2 ? 288 : 144;
    if ((!out_jac && !out_keep_dev) || (n && !scalars)) return fail(DG_ERR_BAD_ARG, "msm: null pointer");
    if ((handle != 0) == (bases != nullptr) && n) return fail(DG_ERR_BAD_ARG, "msm: give exactly one of bases_handle / bases");
    ThreadState &t = tls();
    const void *bases_dev = nullptr;
    MsmPre pre = {0, 0, 0};
    if (handle) {
        HandleRec rec;
        rc = lookup_bases<G2>(handle, n, "msm", rec);
        if (rc) return rc;
        bases_dev = rec.dev;
        pre = msm_pre_of(rec);
    }
    size_t need = (G2 ? msm_scratch_bytes_g2(n, pre) : msm_scratch_bytes_g1(n, pre)) + Arena::pad(32 * n) + Arena::pad(JAC) +
                  (handle ? 0 : Arena::pad(PT * n));
    rc = t.arena.ensure(need, t.stream);
    if (rc) return rc;
    uint8_t *d_scalars = t.arena.alloc<uint8_t>(32 * n);
    uint8_t *d_out = t.arena.alloc<uint8_t>(JAC);
    if (out_keep_dev) d_out = (uint8_t *)out_keep_dev;
    // Large inputs arrive in four chunks on a second stream so the PCIe transfer overlaps the histogram pass (pinned
    // host memory makes the copies truly asynchronous; pageable memory still works, without the overlap).
    MsmStage stage = {1, {}, {0, n, n, n, n}};
    const bool chunked = n >= (1u << 18) && !mont_scalars && ctx().tunable[7].load() == 0;      // tunable 7: A/B switch
    if (chunked) {
        if (!t.stream2) {
            DG_CUDA(cudaStreamCreateWithFlags(&t.stream2, cudaStreamNonBlocking));
            DG_CUDA(cudaEventCreateWithFlags(&t.ev_a, cudaEventDisableTiming));
            DG_CUDA(cudaEventCreateWithFlags(&t.ev_b, cudaEventDisableTiming));
        }
        if (!t.stage_ev[0])
            for (int k = 0; k < 4; k++) DG_CUDA(cudaEventCreateWithFlags(&t.stage_ev[k], cudaEventDisableTiming));
        DG_CUDA(cudaEventRecord(t.ev_a, t.stream));         // the scratch may still be read by work queued on t.stream
        DG_CUDA(cudaStreamWaitEvent(t.stream2, t.ev_a, 0));
        stage.nchunks = 4;
        for (int k = 0; k < 4; k++) {
            stage.lo[k] = n * k / 4;
            stage.lo[k + 1] = n * (k + 1) / 4;
            stage.ev[k] = t.stage_ev[k];
            const size_t lo = stage.lo[k], cnt = stage.lo[k + 1] - lo;
            DG_CUDA(cudaMemcpyAsync(d_scalars + 32 * lo, scalars + 32 * lo, 32 * cnt, cudaMemcpyHostToDevice, t.stream2));
            DG_CUDA(cudaEventRecord(stage.ev[k], t.stream2));
        }
    } else if (n) {
        DG_CUDA(cudaMemcpyAsync(d_scalars, scalars, 32 * n, cudaMemcpyHostToDevice, t.stream));
    }
    if (n && mont_scalars) fr_into_bigint_device(d_scalars, d_scalars, n, t.stream);   // msm_unchecked: into_bigint first
    if (!handle && n) {
        uint8_t *d_bases = t.arena.alloc<uint8_t>(PT * n);
        DG_CUDA(cudaMemcpyAsync(d_bases, bases, PT * n, cudaMemcpyHostToDevice, t.stream));
        bases_dev = d_bases;
    }
    char *scratch = t.arena.alloc<char>(need - t.arena.used);
    rc = G2 ? msm_run_g2(bases_dev, d_scalars, n, d_out, scratch, t.err_flag, t.stream, pre, chunked ? &stage : nullptr)
            : msm_run_g1(bases_dev, d_scalars, n, d_out, scratch, t.err_flag, t.stream, pre, chunked ? &stage : nullptr);
    if (rc) return rc;
    if (!out_keep_dev) DG_CUDA(cudaMemcpyAsync(out_jac, d_out, JAC, cudaMemcpyDeviceToHost, t.stream));
    return read_err_flag(t, t.stream, "msm");
}

// Asynchronous: returns once the kernels are queued on `stream`.  Scalars are NOT validated here (a scalar >= r is
// dropped by k_digits and raises the thread's device-side flag): dg_stream_status reports it after the fact.
template <bool G2>
static int32_t msm_device(uint64_t handle, const void *bases_dev, const void *scalars_dev, size_t n, void *out_jac_dev, void *stream) {
    int32_t rc = check_init();
    if (rc) return rc;
    MsmPre pre = {0, 0, 0};
    if (handle) {
        HandleRec rec;
        rc = lookup_bases<G2>(handle, n, "msm_device", rec);
        if (rc) return rc;
        bases_dev = rec.dev;
        pre = msm_pre_of(rec);
    }
    if (!out_jac_dev || (n && (!bases_dev || !scalars_dev))) return fail(DG_ERR_BAD_ARG, "msm_device: null pointer");
    ThreadState &t = tls();
    cudaStream_t s = stream ? (cudaStream_t)stream : t.stream;
    size_t need = G2 ? msm_scratch_bytes_g2(n, pre) : msm_scratch_bytes_g1(n, pre);
    rc = t.arena.ensure(need, s);
    if (rc) return rc;
    char *scratch = t.arena.alloc<char>(need);
    rc = G2 ? msm_run_g2(bases_dev, scalars_dev, n, out_jac_dev, scratch, t.err_flag, s, pre)
            : msm_run_g1(bases_dev, scalars_dev, n, out_jac_dev, scratch, t.err_flag, s, pre);
    if (rc) return rc;
    return t.arena.mark(s);
}

template <bool G2> static int32_t bases_upload(const uint8_t *affine, size_t n, uint64_t *handle) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!affine || !handle || n == 0) return fail(DG_ERR_BAD_ARG, "bases_upload: null pointer or n == 0");
    const size_t PT = G2 ? 192 : 96;
    if (n >= (1ull << 30)) return fail(DG_ERR_BAD_ARG, "bases_upload: n must be < 2^30");
    // The resident array holds the points followed by their GLV images phi(P_i) = (beta x_i, y_i) (one field
    // multiplication per point, done here once instead of inside every MSM).
    void *dev = nullptr;
    DG_CUDA(cudaMalloc(&dev, 2 * PT * n));
    ThreadState &t = tls();
    cudaError_t e = cudaMemcpyAsync(dev, affine, PT * n, cudaMemcpyHostToDevice, t.stream);
    if (e == cudaSuccess && (G2 ? glv_expand_g2(dev, n, dev, n, t.stream) : glv_expand_g1(dev, n, dev, n, t.stream)) != DG_OK) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaStreamSynchronize(t.stream);
    if (e != cudaSuccess) { cudaFree(dev); return fail(DG_ERR_CUDA, cudaGetErrorString(e)); }
    std::lock_guard<std::mutex> lk(ctx().mu);
    uint64_t h = ctx().next_handle++;
    HandleRec r;
    r.kind = G2 ? HandleRec::BASES_G2 : HandleRec::BASES_G1;
    r.dev = dev; r.n = n; r.slot = t.slot; r.phi_off = n;
    ctx().handles[h] = r;
    *handle = h;
    return DG_OK;
}

// dg_bases_precompute / dg_bases_free for one single-device handle, on the calling thread's device
static int32_t precompute_one(uint64_t handle, int32_t c) {
    int32_t rc = check_init();
    if (rc) return rc;
    HandleRec rec;
    {
        std::lock_guard<std::mutex> lk(ctx().mu);
        auto it = ctx().handles.find(handle);
        if (it == ctx().handles.end() || (it->second.kind != HandleRec::BASES_G1 && it->second.kind != HandleRec::BASES_G2))
            return fail(DG_ERR_BAD_ARG, "bases_precompute: bad handle");
        if (it->second.slot != tls().slot) return fail(DG_ERR_BAD_ARG, "bases_precompute: the handle lives on another device");
        rec = it->second;
    }
    if (c == 0) c = rec.n >= (1u << 22) ? 20 : 17;     // 13 / 15 rows; measured with the batch-affine stage (tools/sweep_rounds.py)
    DG_CUDA(cudaDeviceSynchronize());
    rc = rec.kind == HandleRec::BASES_G2 ? bases_precompute_g2(rec, c, tls().stream) : bases_precompute_g1(rec, c, tls().stream);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx().mu);
    ctx().handles[handle] = rec;                        // rec.dev now points at the table (the plain copy was freed)
    return DG_OK;
}
static int32_t free_one(uint64_t handle) {
    void *dev = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx().mu);
        auto it = ctx().handles.find(handle);
        if (it == ctx().handles.end() || (it->second.kind != HandleRec::BASES_G1 && it->second.kind != HandleRec::BASES_G2))
            return fail(DG_ERR_BAD_ARG, "bases_free: bad handle");
        if (it->second.slot != tls().slot) return fail(DG_ERR_BAD_ARG, "bases_free: the handle lives on another device");
        dev = it->second.dev;
        ctx().handles.erase(it);
    }
    cudaDeviceSynchronize();
    cudaFree(dev);
    return DG_OK;
}

// ---- single-process multi-GPU (SURVEY.md 8b "dg_msm_g1_sharded", 8e) --------------------------------------------------
// Contiguous base ranges, sizes differing by at most one (the same split crypto_b200/sharding.py::shard_range makes).
static void shard_bounds(size_t n, int ndev, std::vector<size_t> &lo) {
    lo.assign(ndev + 1, 0);
    size_t base = n / ndev, rem = n % ndev;
    for (int d = 0; d < ndev; d++) lo[d + 1] = lo[d] + base + ((size_t)d < rem ? 1 : 0);
}
static int32_t lookup_sharded(uint64_t handle, bool g2, HandleRec &out) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || it->second.kind != (g2 ? HandleRec::SHARDED_G2 : HandleRec::SHARDED_G1))
        return fail(DG_ERR_BAD_ARG, "bad sharded bases handle");
    out = it->second;
    return DG_OK;
}

template <bool G2> static int32_t bases_upload_sharded(const uint8_t *affine, size_t n, uint64_t *handle) {
    int32_t rc = check_init();
    if (rc) return rc;
    if (!affine || !handle || n == 0) return fail(DG_ERR_BAD_ARG, "bases_upload_sharded: null pointer or n == 0");
    const size_t PT = G2 ? 192 : 96;
    const int ndev = ctx().ndev;
    HandleRec r;
    r.kind = G2 ? HandleRec::SHARDED_G2 : HandleRec::SHARDED_G1;
    r.n = n;
    shard_bounds(n, ndev, r.shard_lo);
    r.shard_handle.assign(ndev, 0);
    rc = run_on_devices(ndev, [&](int d) -> int32_t {
        size_t lo = r.shard_lo[d], cnt = r.shard_lo[d + 1] - lo;
        if (!cnt) return DG_OK;                          // fewer points than devices
        return bases_upload<G2>(affine + PT * lo, cnt, &r.shard_handle[d]);
    });
    if (rc) {
        std::string keep = tls().err;
        run_on_devices(ndev, [&](int d) -> int32_t { return r.shard_handle[d] ? free_one(r.shard_handle[d]) : DG_OK; });
        return fail(rc, keep);
    }
    std::lock_guard<std::mutex> lk(ctx().mu);
    uint64_t h = ctx().next_handle++;
    ctx().handles[h] = r;
    *handle = h;
    return DG_OK;
}

template <bool G2>
static int32_t msm_sharded(uint64_t handle, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac, bool mont_scalars) {
    int32_t rc = check_init();
    if (rc) return rc;
    const size_t PT = G2 ? 192 : 96, JAC = G2 ? 288 : 144;
    if (!out_jac || (n && !scalars)) return fail(DG_ERR_BAD_ARG, "msm_sharded: null pointer");
    if ((handle != 0) == (bases != nullptr) && n) return fail(DG_ERR_BAD_ARG, "msm_sharded: give exactly one of bases_handle / bases");
    Context &c = ctx();
    const int ndev = c.ndev;
    HandleRec rec;
    if (handle) {
        rc = lookup_sharded(handle, G2, rec);
        if (rc) return rc;
        if (n > rec.n) return fail(DG_ERR_BAD_ARG, "msm_sharded: n exceeds uploaded bases");
    } else {
        shard_bounds(n, ndev, rec.shard_lo);
        rec.shard_handle.assign(ndev, 0);
    }
    // one sharded MSM at a time: the per-device partial buffers are shared (the call occupies every device anyway)
    std::lock_guard<std::mutex> big(c.sharded_mu);
    rc = run_on_devices(ndev, [&](int d) -> int32_t {
        Worker *w = ctx().workers[d];
        if (!w->partial) DG_CUDA(cudaMalloc(&w->partial, 512));
        size_t lo = rec.shard_lo[d], hi = rec.shard_lo[d + 1] < n ? rec.shard_lo[d + 1] : n;
        size_t cnt = hi > lo ? hi - lo : 0;
        return msm_host<G2>(cnt ? rec.shard_handle[d] : 0, (cnt && !handle) ? bases + PT * lo : nullptr, scalars + 32 * lo, cnt, nullptr,
                            mont_scalars, w->partial);
    });
    if (rc) return rc;
    // every device's stream is idle again (msm_host synchronised): fold the partials on device 0, reading them
    // straight out of the peers' memory over NVLink where peer access is available
    return run_on_devices(1, [&](int) -> int32_t {
        ThreadState &t = tls();
        int32_t rc2 = t.arena.ensure(Arena::pad(JAC * (ndev + 1)), t.stream);
        if (rc2) return rc2;
        uint8_t *d_out = t.arena.alloc<uint8_t>(JAC), *d_stage = t.arena.alloc<uint8_t>(JAC * ndev);
        PtrList pl = {};
        for (int d = 0; d < ndev; d++) {
            if (d == 0 || ctx().peer[d]) { pl.p[d] = ctx().workers[d]->partial; continue; }
            DG_CUDA(cudaMemcpyPeerAsync(d_stage + JAC * d, ctx().devices[0], ctx().workers[d]->partial, ctx().devices[d], JAC, t.stream));
            pl.p[d] = d_stage + JAC * d;
        }
        rc2 = G2 ? fold_ptrs_g2(pl, ndev, d_out, t.stream) : fold_ptrs_g1(pl, ndev, d_out, t.stream);
        if (rc2) return rc2;
        DG_CUDA(cudaMemcpyAsync(out_jac, d_out, JAC, cudaMemcpyDeviceToHost, t.stream));
        DG_CUDA(cudaStreamSynchronize(t.stream));
        return DG_OK;
    });
}

}  // namespace dg

using namespace dg;

extern "C" {

int32_t dg_init_devices(const int32_t *devices, int32_t ndev) {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (!devices || ndev < 1 || ndev > DG_MAX_DEVICES) return fail(DG_ERR_BAD_ARG, "dg_init_devices: need 1..16 devices");
    if (c.inited) {
        bool same = ndev == c.ndev;
        for (int d = 0; same && d < ndev; d++) same = devices[d] == c.devices[d] || (ndev == 1 && devices[d] < 0);
        if (same || ndev == 1) return DG_OK;             // idempotent; dg_init on a multi-device context is a no-op
        return fail(DG_ERR_BAD_ARG, "dg_init_devices: already initialised with a different device list (dg_shutdown first)");
    }
    int count = 0;
    DG_CUDA(cudaGetDeviceCount(&count));
    if (count == 0) return fail(DG_ERR_CUDA, "no CUDA device visible: libdockgpu has no CPU fallback");
    int devs[DG_MAX_DEVICES];
    for (int d = 0; d < ndev; d++) {
        devs[d] = devices[d];
        if (devs[d] < 0) {
            if (ndev != 1) return fail(DG_ERR_BAD_ARG, "dg_init_devices: negative device index");
            DG_CUDA(cudaGetDevice(&devs[d]));
        }
        if (devs[d] >= count) return fail(DG_ERR_BAD_ARG, "dg_init: device index out of range");
        for (int e = 0; e < d; e++)
            if (devs[e] == devs[d]) return fail(DG_ERR_BAD_ARG, "dg_init_devices: duplicate device");
    }
    int sm_count = 0;
    for (int d = 0; d < ndev; d++) {
        cudaDeviceProp prop;
        DG_CUDA(cudaGetDeviceProperties(&prop, devs[d]));
        if (prop.major < 10) return fail(DG_ERR_CUDA, "libdockgpu is built for sm_100a (B200) only");
        if (d == 0) sm_count = prop.multiProcessorCount;
        DG_CUDA(cudaSetDevice(devs[d]));
        // The hot kernels gather 96-byte records at random from tables far larger than L2; ask for the
        // smallest DRAM->L2 fetch granularity so a miss does not drag the whole 128-byte line in.
        size_t gran = 32;
        if (const char *e = getenv("DG_L2_FETCH_GRANULARITY")) gran = (size_t)atoi(e);
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    // device 0 folds the partial results of a sharded MSM out of its peers' memory
    DG_CUDA(cudaSetDevice(devs[0]));
    c.peer[0] = true;
    for (int d = 1; d < ndev; d++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devs[0], devs[d]);
        if (can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[d], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
            cudaGetLastError();
        }
        c.peer[d] = can != 0;
    }
    for (int d = 0; d < ndev; d++) c.devices[d] = devs[d];
    c.ndev = ndev;
    c.device = devs[0];
    c.sm_count = sm_count;
    c.inited = true;
    return DG_OK;
}

int32_t dg_init(int32_t device) { return dg_init_devices(&device, 1); }

int32_t dg_device_count(int32_t *ndev) {
    if (!ndev) return fail(DG_ERR_BAD_ARG, "dg_device_count: null pointer");
    if (!ctx().inited) return fail(DG_ERR_NOT_INIT, "dg_init has not been called");
    *ndev = ctx().ndev;
    return DG_OK;
}

int32_t dg_shutdown(void) {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.inited) return DG_OK;
    stop_workers();
    for (int d = 0; d < c.ndev; d++) {
        cudaSetDevice(c.devices[d]);
        cudaDeviceSynchronize();
    }
    cudaSetDevice(c.devices[0]);
    for (auto &kv : c.handles) cudaFree(kv.second.dev);
    c.handles.clear();
    ntt_release_plans();
    ThreadState &t = tls();
    t.arena.release();
    {
        std::lock_guard<std::mutex> lk2(c.attr_mu);
        c.attr_done.clear();
    }
    c.inited = false;
    c.ndev = 0;
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

int32_t dg_stream_status(void *stream) {
    int32_t rc = check_init();
    if (rc) return rc;
    ThreadState &t = tls();
    cudaStream_t s = stream ? (cudaStream_t)stream : t.stream;
    return read_err_flag(t, s, "msm_device");
}

int32_t dg_bases_upload_g1(const uint8_t *affine, size_t n, uint64_t *handle) { return bases_upload<false>(affine, n, handle); }
int32_t dg_bases_upload_g2(const uint8_t *affine, size_t n, uint64_t *handle) { return bases_upload<true>(affine, n, handle); }
int32_t dg_bases_upload_g1_sharded(const uint8_t *affine, size_t n, uint64_t *handle) { return bases_upload_sharded<false>(affine, n, handle); }
int32_t dg_bases_upload_g2_sharded(const uint8_t *affine, size_t n, uint64_t *handle) { return bases_upload_sharded<true>(affine, n, handle); }

static bool is_sharded(uint64_t handle, HandleRec &rec) {
    std::lock_guard<std::mutex> lk(ctx().mu);
    auto it = ctx().handles.find(handle);
    if (it == ctx().handles.end() || (it->second.kind != HandleRec::SHARDED_G1 && it->second.kind != HandleRec::SHARDED_G2)) return false;
    rec = it->second;
    return true;
}

int32_t dg_bases_free(uint64_t handle) {
    HandleRec rec;
    if (is_sharded(handle, rec)) {
        int32_t rc = run_on_devices(ctx().ndev, [&](int d) -> int32_t { return rec.shard_handle[d] ? free_one(rec.shard_handle[d]) : DG_OK; });
        std::lock_guard<std::mutex> lk(ctx().mu);
        ctx().handles.erase(handle);
        return rc;
    }
    return free_one(handle);
}

int32_t dg_msm_g1(uint64_t h, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out) { return msm_host<false>(h, bases, scalars, n, out); }
int32_t dg_msm_g2(uint64_t h, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out) { return msm_host<true>(h, bases, scalars, n, out); }
int32_t dg_msm_unchecked_g1(uint64_t h, const uint8_t *bases, const uint8_t *fr_mont, size_t n, uint8_t *out) { return msm_host<false>(h, bases, fr_mont, n, out, true); }
int32_t dg_msm_unchecked_g2(uint64_t h, const uint8_t *bases, const uint8_t *fr_mont, size_t n, uint8_t *out) { return msm_host<true>(h, bases, fr_mont, n, out, true); }
int32_t dg_msm_g1_sharded(uint64_t h, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out) { return msm_sharded<false>(h, bases, scalars, n, out, false); }
int32_t dg_msm_g2_sharded(uint64_t h, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out) { return msm_sharded<true>(h, bases, scalars, n, out, false); }
int32_t dg_msm_unchecked_g1_sharded(uint64_t h, const uint8_t *bases, const uint8_t *fr_mont, size_t n, uint8_t *out) { return msm_sharded<false>(h, bases, fr_mont, n, out, true); }
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
    HandleRec rec;
    if (is_sharded(handle, rec))                        // every device builds the table of its own base range
        return run_on_devices(ctx().ndev, [&](int d) -> int32_t { return rec.shard_handle[d] ? precompute_one(rec.shard_handle[d], c) : DG_OK; });
    return precompute_one(handle, c);
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
    MsmPre pre = {precomputed_window_bits, precomputed_window_bits ? (uint32_t)n : 0u, 0u};
    int c = 0, r = 0;
    if (is_g2) msm_plan_g2(n, pre, &c, &r); else msm_plan_g1(n, pre, &c, &r);
    *window_bits = c;
    *affine_rounds = r;
    return DG_OK;
}

}  // extern "C"
