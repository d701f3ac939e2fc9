"""G2 MSM geometry sweep: python tools/sweep_g2.py LOGN c1,c2,... [r1,r2,...] [w1,w2..] [pre]
(c = 0: automatic; rounds -1: automatic; w = tunable 6: 0 automatic, 1 4-lane quads, 2 12-lane quads for the line sums;
pre: resident 2^(ck)-multiples table instead of raw bases + GLV, then c is the table window, 0 = default)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
logn = int(sys.argv[1]); cs = [int(x) for x in sys.argv[2].split(',')]
rs = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [-1]
ws = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0]
pre = len(sys.argv) > 5 and sys.argv[5] == 'pre'
n = 1 << logn
sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
lib.init()
one = np.zeros(32, np.uint8); one[0] = 1
tbl = lib.FixedBaseTable(cref.g2_generator_muls(one), n, g2=True)
bases = np.array(tbl.mul_many_normalized(ks))
dot = np.frombuffer(cref.scalar_dot_mod_r(ks, sc).to_bytes(32, 'little'), dtype=np.uint8)
exp = bytes(np.array(tbl.mul_many_normalized(dot)))
tbl.free()
hb = lib.Bases(bases, g2=True)
d_s = torch.from_numpy(np.array(sc)).cuda(); d_o = torch.zeros(288, dtype=torch.uint8, device='cuda')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for c, r, w in [(c, r, w) for c in cs for r in rs for w in ws]:
    if True:
        if pre:
            hb.free(); hb = lib.Bases(bases, g2=True).precompute(c)
        else:
            lib.msm_set_window(c)
        lib.msm_set_affine_rounds(r); lib.dbg_set_tunable(6, w)
        for _ in range(2):
            lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
        torch.cuda.synchronize()
        ok = bytes(lib.normalize_batch(d_o.cpu().numpy(), g2=True)) == exp
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print('G2 2^%d %s c=%2d rounds=%2d quads=%d: %8.3f ms  ok=%s' % (logn, 'table' if pre else 'raw', c, r, w, ms, ok), flush=True)
