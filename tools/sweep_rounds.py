"""G1 (or, with env G2=1, G2) MSM timing over (precompute window, batch-affine rounds):
python tools/sweep_rounds.py LOGN c1,c2,.. r1,r2,.. [MIN_WAVES [_ [KMAX]]]   (r = -1: automatic rounds)
c = 0 means plain bases (window from the heuristic, or `cW` entries like 0:14 to force window 14)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib
logn = int(sys.argv[1])
cs = sys.argv[2].split(',')
rs = [int(x) for x in sys.argv[3].split(',')]
n = 1 << logn
sc = cref.random_scalars(n, 300 + logn); ks = cref.random_scalars(n, 400 + logn)
G2 = os.environ.get('G2') == '1'
bases = cref.g2_generator_muls(ks) if G2 else cref.g1_generator_muls(ks)
lib.init()
if len(sys.argv) > 4:
    lib.dbg_set_tunable(0, int(sys.argv[4]))
if len(sys.argv) > 6:
    lib.dbg_set_tunable(3, int(sys.argv[6]))
d_s = torch.from_numpy(sc).cuda(); d_o = torch.zeros(288 if G2 else 144, dtype=torch.uint8, device='cuda')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
tot = 0
for a, b in zip(np.asarray(ks).reshape(-1, 32), np.asarray(sc).reshape(-1, 32)):
    tot += int.from_bytes(bytes(a), 'little') * int.from_bytes(bytes(b), 'little')
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
exp = bytes((cref.g2_generator_muls if G2 else cref.g1_generator_muls)(np.frombuffer((tot % R).to_bytes(32, 'little'), dtype=np.uint8)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for cspec in cs:
    c, _, w = cspec.partition(':')
    c = int(c); w = int(w) if w else 0
    hb = lib.Bases(bases, g2=G2)
    if c:
        hb.precompute(c)
    lib.msm_set_window(w)
    for r in rs:
        lib.msm_set_affine_rounds(r)
        for _ in range(2):
            lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
        torch.cuda.synchronize()
        ok = bytes((cref.normalize_batch_g2 if G2 else cref.normalize_batch_g1)(d_o.cpu().numpy())) == exp
        best = 1e9; tot_ms = 0
        for _ in range(5):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); best = min(best, ms); tot_ms += ms
        print('2^%d pre c=%2d win=%2d rounds=%2d: mean %8.3f ms best %8.3f ms  %.3e muls/s  ok=%s' %
              (logn, c, w, r, tot_ms / 5, best, n / (tot_ms / 5) * 1e3, ok), flush=True)
    lib.msm_set_window(0)
    lib.msm_set_affine_rounds(-1)
    hb.free()
