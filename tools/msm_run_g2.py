"""Run a G2 MSM twice after warm-up (for ncu): python tools/msm_run_g2.py LOGN C [W]
(C = 0: plain handle, else table window; W: bucket window override for the plain handle)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
logn = int(sys.argv[1]); c = int(sys.argv[2])
n = 1 << logn
sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
lib.init()
one = np.zeros(32, np.uint8); one[0] = 1
tbl = lib.FixedBaseTable(cref.g2_generator_muls(one), n, g2=True)
bases = np.array(tbl.mul_many_normalized(ks)); tbl.free()
hb = lib.Bases(bases, g2=True)
if c:
    hb.precompute(c)
d_s = torch.from_numpy(np.array(sc)).cuda(); d_o = torch.zeros(288, dtype=torch.uint8, device='cuda')
if len(sys.argv) > 3:
    lib.msm_set_window(int(sys.argv[3]))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for _ in range(2):
    lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
torch.cuda.synchronize()
