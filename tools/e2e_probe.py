"""Where the host-call overhead of dg_msm_g1 goes (measurement script): device-timed MSM, bare H2D of the scalars, the
host call with and without chunked staging."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import cref
from crypto_b200 import lib
lib.init(0)
n = 1 << 20
ss = cref.random_scalars(n, 1); ks = cref.random_scalars(n, 2)
import bench
bases = bench.gpu_bases(lib, ks)
pin = torch.from_numpy(np.array(ss)).pin_memory(); pin_np = pin.numpy()
d_s = torch.from_numpy(np.array(ss)).cuda(); d_o = torch.zeros(144, dtype=torch.uint8, device='cuda')
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for pre in (False, True):
    hb = lib.Bases(bases)
    if pre:
        hb.precompute(0)
    for _ in range(3):
        lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    dev_ms = (time.perf_counter() - t) * 100
    t = time.perf_counter()
    for _ in range(10):
        d_s.copy_(pin, non_blocking=True)
    torch.cuda.synchronize()
    h2d_ms = (time.perf_counter() - t) * 100
    res = {}
    for chunk in (0, 1):
        lib.dbg_set_tunable(7, chunk)
        for _ in range(3):
            lib.msm(hb, pin_np)
        t = time.perf_counter()
        for _ in range(10):
            lib.msm(hb, pin_np)
        res['host_call_ms_%s' % ('single_copy' if chunk else 'chunked')] = (time.perf_counter() - t) * 100
    lib.dbg_set_tunable(7, 0)
    print('table' if pre else 'plain', 'device %.3f ms, H2D alone %.3f ms' % (dev_ms, h2d_ms), res, flush=True)
    hb.free()
