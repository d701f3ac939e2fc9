"""G2 MSM timing at 2^16 / 2^18 / 2^20 terms, raw bases and resident table, each checked against the known-dlog identity;
DOCKGPU_LIB selects an alternative build for A/B runs: python tools/g2_ab.py"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
lib.init()
one = np.zeros(32, np.uint8); one[0] = 1
for logn in (16, 18, 20):
    n = 1 << logn
    sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
    tbl = lib.FixedBaseTable(cref.g2_generator_muls(one), n, g2=True)
    bases = np.array(tbl.mul_many_normalized(ks))
    dot = np.frombuffer(cref.scalar_dot_mod_r(ks, sc).to_bytes(32, 'little'), dtype=np.uint8)
    exp = bytes(np.array(tbl.mul_many_normalized(dot))); tbl.free()
    hb = lib.Bases(bases, g2=True)
    d_s = torch.from_numpy(np.array(sc)).cuda(); d_o = torch.zeros(288, dtype=torch.uint8, device='cuda')
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    res = []
    for mode in ('raw', 'table'):
        if mode == 'table': hb.precompute()
        for _ in range(2): lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
        torch.cuda.synchronize()
        ok = bytes(lib.normalize_batch(d_o.cpu().numpy(), g2=True)) == exp
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        res.append('%s %.3f ms ok=%s' % (mode, e0.elapsed_time(e1) / 4, ok))
    hb.free()
    print('G2 2^%d:' % logn, ' | '.join(res), flush=True)
