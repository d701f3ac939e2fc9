"""Run the G1 MSM a few times on device-resident synthetic inputs (for ncu / timing).
usage: python tools/msm_run.py LOGN [C] [ITERS]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import cref
from crypto_b200 import lib

logn = int(sys.argv[1]); c = int(sys.argv[2]) if len(sys.argv) > 2 else 0; iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = 1 << logn
sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
bases = cref.g1_generator_muls(ks)
lib.init()
lib.msm_set_window(c)
d_b = torch.from_numpy(bases).cuda(); d_s = torch.from_numpy(sc).cuda(); d_o = torch.zeros(144, dtype=torch.uint8, device='cuda')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for _ in range(iters):
    lib.msm_device(d_b.data_ptr(), d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    lib.msm_device(d_b.data_ptr(), d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
e1.record(); torch.cuda.synchronize()
print('msm g1 2^%d c=%d: %.3f ms/iter' % (logn, c, e0.elapsed_time(e1) / iters))
