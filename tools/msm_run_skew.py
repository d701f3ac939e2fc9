"""Witness-shaped (skewed) scalars through a precomputed handle, for ncu launch lists:
   python tools/msm_run_skew.py LOGN   (10 % zeros, 10 % ones, 10 % 16-bit values, rest uniform -- tools/replay_workloads.py config 3)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
logn = int(sys.argv[1])
n = 1 << logn
sc = cref.random_scalars(n, 300 + logn).copy(); ks = cref.random_scalars(n, 400 + logn)
sv = sc.reshape(n, 32)
sv[: n // 10] = 0
sv[n // 10: n // 5] = 0
sv[n // 10: n // 5, 0] = 1
sv[n // 5: 3 * n // 10, 2:] = 0
bases = cref.g1_generator_muls(ks)
lib.init()
hb = lib.Bases(bases).precompute()
d_s = torch.from_numpy(sc).cuda(); d_o = torch.zeros(144, dtype=torch.uint8, device='cuda')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for _ in range(3):
    lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(ts)
lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
e1.record(ts); torch.cuda.synchronize()
print('skewed 2^%d: %.3f ms' % (logn, e0.elapsed_time(e1)))
