"""python tools/pairing_run.py K  -- run a K-pair multi_pairing twice (for ncu launch lists)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cref
from crypto_b200 import lib
k = int(sys.argv[1])
ps = cref.g1_generator_muls(cref.random_scalars(k, 41)); qs = cref.g2_generator_muls(cref.random_scalars(k, 42))
lib.init()
lib.multi_pairing(ps, qs)
t = time.perf_counter(); out = lib.multi_pairing(ps, qs); dt = time.perf_counter() - t
print('k=%d host call %.3f ms ok=%s' % (k, dt * 1e3, bytes(out) == bytes(cref.multi_pairing(ps, qs))))
