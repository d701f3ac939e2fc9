"""A/B of the window-group split (tunable 2; crypto_b200/csrc/msm_host.cuh msm_run): python tools/split_ab.py g1|g2 LOGN[,LOGN..] [h1,h2,..] [kmax1,..]
(h = 0: no split, h >= 2: h high windows on the second stream; kmax = tunable 3, 0 = default).  Raw bases behind a plain handle, device
scalars, the known-dlog identity checks every result."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
g2 = sys.argv[1] == 'g2'
logns = [int(x) for x in sys.argv[2].split(',')]
hs = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 4]
kmaxs = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0]      # tunable 3: outputs per thread of a batch-affine round
lib.init()
one = np.zeros(32, np.uint8); one[0] = 1
JAC = 288 if g2 else 144
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for logn in logns:
    n = 1 << logn
    sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
    gen = cref.g2_generator_muls(one) if g2 else cref.g1_generator_muls(one)
    tbl = lib.FixedBaseTable(gen, n, g2=g2)
    bases = np.array(tbl.mul_many_normalized(ks))
    dot = np.frombuffer(cref.scalar_dot_mod_r(ks, sc).to_bytes(32, 'little'), dtype=np.uint8)
    exp = bytes(np.array(tbl.mul_many_normalized(dot)))
    tbl.free()
    hb = lib.Bases(bases, g2=g2)
    d_s = torch.from_numpy(np.array(sc)).cuda(); d_o = torch.zeros(JAC, dtype=torch.uint8, device='cuda')
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    for h, km in [(h, km) for km in kmaxs for h in hs]:
        lib.dbg_set_tunable(2, h); lib.dbg_set_tunable(3, km)
        for _ in range(3):
            lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
        torch.cuda.synchronize()
        ok = bytes(lib.normalize_batch(d_o.cpu().numpy(), g2=g2)) == exp
        tot = 0.0
        reps = 5
        for _ in range(reps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
            e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        print('%s 2^%d split=%d kmax=%d: %8.3f ms  ok=%s' % ('G2' if g2 else 'G1', logn, h, km, tot / reps, ok), flush=True)
    lib.dbg_set_tunable(2, 0); lib.dbg_set_tunable(3, 0)
    hb.free()
