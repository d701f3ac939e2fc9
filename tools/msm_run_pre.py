"""Run the precomputed-bases G1 MSM once more after warm-up (for ncu): python tools/msm_run_pre.py LOGN C [ROUNDS]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
logn = int(sys.argv[1]); c = int(sys.argv[2])
n = 1 << logn
sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
bases = cref.g1_generator_muls(ks)
lib.init()
if len(sys.argv) > 3:
    lib.msm_set_affine_rounds(int(sys.argv[3]))
hb = lib.Bases(bases)
if c:
    hb.precompute(c)
d_s = torch.from_numpy(sc).cuda(); d_o = torch.zeros(144, dtype=torch.uint8, device='cuda')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for _ in range(2):
    lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
torch.cuda.synchronize()
