"""Timing of dg_multi_pairing / dg_multi_pairing_batch against the pair count (measurement script)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cref
from crypto_b200 import lib
lib.init()
K = 4096
ps = cref.g1_generator_muls(cref.random_scalars(K, 41)); qs = cref.g2_generator_muls(cref.random_scalars(K, 42))
def best(fn, reps=3):
    fn(); b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t)
    return b * 1e3
for k in (64, 256, 512, 640, 1024, 1280, 2048, 4096):
    print('multi_pairing(%d): %.3f ms' % (k, best(lambda: lib.multi_pairing(ps[:96 * k], qs[:192 * k]))), flush=True)
for nb, m in ((10, 64), (10, 128), (6, 128), (4, 256)):
    print('multi_pairing_batch(%d x %d): %.3f ms' % (nb, m, best(lambda: lib.multi_pairing_batch(ps[:96 * nb * m], qs[:192 * nb * m], [m] * nb))), flush=True)
