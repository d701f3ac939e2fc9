import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import cref
L = C.CDLL('/root/repo/crypto_b200/libdockgpu.so')
def chk(rc):
    if rc != 0:
        buf = C.create_string_buffer(512); L.dg_last_error(buf, C.c_size_t(512)); raise RuntimeError((rc, buf.value))
chk(L.dg_init(C.c_int32(-1)))
vp = lambda a: C.c_void_p(a.ctypes.data)
# field ops
n = 4096
a = cref.g1_generator_muls(cref.random_scalars(n, 11)).reshape(n, 96)[:, :48].copy().reshape(-1)
b = cref.g1_generator_muls(cref.random_scalars(n, 12)).reshape(n, 96)[:, 48:].copy().reshape(-1)
out = np.zeros(48 * n, np.uint8)
chk(L.dg_dbg_fp_op(C.c_int32(0), vp(a), vp(b), C.c_size_t(n), vp(out)))
exp = np.zeros(48 * n, np.uint8)
for i in range(n):
    cref.lib().ref_fp_mul(vp(a[48*i:]), vp(b[48*i:]), vp(exp[48*i:]))
print('fp_mul parity', np.array_equal(out, exp))
chk(L.dg_dbg_fp_op(C.c_int32(5), vp(a), vp(b), C.c_size_t(n), vp(out)))
for i in range(16):
    cref.lib().ref_fp_inv(vp(a[48*i:]), vp(exp[48*i:]))
print('fp_inv parity', np.array_equal(out[:48*16], exp[:48*16]))
# msm
for logn in (4, 16):
    n = 1 << logn
    sc = cref.random_scalars(n, 100 + logn); ks = cref.random_scalars(n, 200 + logn)
    bases = cref.g1_generator_muls(ks)
    out = np.zeros(144, np.uint8)
    t = time.time(); chk(L.dg_msm_g1(C.c_uint64(0), vp(bases), vp(sc), C.c_size_t(n), vp(out))); dt = time.time() - t
    t = time.time(); ref = cref.msm_g1(bases, sc); dtc = time.time() - t
    ok = np.array_equal(cref.normalize_batch_g1(out), cref.normalize_batch_g1(ref))
    print('msm g1 2^%d parity %s gpu %.4fs cpu %.4fs' % (logn, ok, dt, dtc))
# timing, device resident
import torch
for logn in (16, 18, 20, 22):
    n = 1 << logn
    sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
    bases = cref.g1_generator_muls(ks)
    d_b = torch.from_numpy(bases).cuda(); d_s = torch.from_numpy(sc).cuda(); d_o = torch.zeros(144, dtype=torch.uint8, device='cuda')
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
    for c in ([0] if logn != 20 else [0, 13, 14, 16, 17]):
        chk(L.dg_msm_set_window(C.c_int32(c)))
        for _ in range(2):
            chk(L.dg_msm_g1_device(C.c_void_p(d_b.data_ptr()), C.c_void_p(d_s.data_ptr()), C.c_size_t(n), C.c_void_p(d_o.data_ptr()), C.c_void_p(st)))
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            chk(L.dg_msm_g1_device(C.c_void_p(d_b.data_ptr()), C.c_void_p(d_s.data_ptr()), C.c_size_t(n), C.c_void_p(d_o.data_ptr()), C.c_void_p(st)))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print('msm g1 2^%d c=%d: %.3f ms  %.3e muls/s' % (logn, c, ms, n / ms * 1e3))
    if logn <= 20:
        ref = cref.msm_g1(bases, sc)
        print('  parity', np.array_equal(cref.normalize_batch_g1(d_o.cpu().numpy()), cref.normalize_batch_g1(ref)))
chk(L.dg_msm_set_window(C.c_int32(0)))
