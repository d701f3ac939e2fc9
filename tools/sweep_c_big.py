"""Window sweep for large raw-bases G1 MSMs (bases made on the GPU): python tools/sweep_c_big.py LOGN c1,c2,... [rounds list]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
logn = int(sys.argv[1]); cs = [int(x) for x in sys.argv[2].split(',')]
rs = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [-1]
n = 1 << logn
sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
lib.init()
one = np.zeros(32, np.uint8); one[0] = 1
tbl = lib.FixedBaseTable(cref.g1_generator_muls(one), n)
bases = np.array(tbl.mul_many_normalized(ks))
dot = np.frombuffer(cref.scalar_dot_mod_r(ks, sc).to_bytes(32, 'little'), dtype=np.uint8)
exp = bytes(np.array(tbl.mul_many_normalized(dot)))
tbl.free()
d_b = torch.from_numpy(bases).cuda(); d_s = torch.from_numpy(np.array(sc)).cuda(); d_o = torch.zeros(144, dtype=torch.uint8, device='cuda')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for c in cs:
    for r in rs:
        lib.msm_set_window(c); lib.msm_set_affine_rounds(r)
        for _ in range(2):
            lib.msm_device(d_b.data_ptr(), d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
        torch.cuda.synchronize()
        ok = bytes(lib.normalize_batch(d_o.cpu().numpy())) == exp
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            lib.msm_device(d_b.data_ptr(), d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        print('G1 2^%d c=%2d rounds=%2d: %8.3f ms  ok=%s' % (logn, c, r, e0.elapsed_time(e1) / 3, ok), flush=True)
