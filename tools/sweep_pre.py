"""Precomputed-bases MSM timing: python tools/sweep_pre.py LOGN c1,c2,..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
logn = int(sys.argv[1]); cs = [int(x) for x in sys.argv[2].split(',')]
n = 1 << logn
sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
bases = cref.g1_generator_muls(ks)
lib.init()
d_s = torch.from_numpy(sc).cuda(); d_o = torch.zeros(144, dtype=torch.uint8, device='cuda')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
tot = 0
for a, b in zip(np.asarray(ks).reshape(-1, 32), np.asarray(sc).reshape(-1, 32)):
    tot += int.from_bytes(bytes(a), 'little') * int.from_bytes(bytes(b), 'little')
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
exp = bytes(cref.g1_generator_muls(np.frombuffer((tot % R).to_bytes(32, 'little'), dtype=np.uint8)))
for c in cs:
    hb = lib.Bases(bases)
    t0 = time.perf_counter()
    if c:
        hb.precompute(c)
    tp = time.perf_counter() - t0
    for _ in range(2):
        lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
    torch.cuda.synchronize()
    ok = bytes(cref.normalize_batch_g1(d_o.cpu().numpy())) == exp
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('2^%d pre c=%2d: %8.3f ms  %.3e muls/s  ok=%s  (precompute %.2f s)' % (logn, c, ms, n / ms * 1e3, ok, tp), flush=True)
    hb.free()
