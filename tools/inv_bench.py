"""Latency of one warp's worth of Fp inversions through the debug hook (measurement script): op 5 Fermat, 6 bit-serial
binary Euclid, 7 Pornin.  The host call overhead (two 1.5 KB copies + one launch) is the same for every op, so the
differences are the kernels' latencies."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from crypto_b200 import lib
from oracle import bls12_381 as o
lib.init(0)
rng = np.random.default_rng(1)
xs = [int.from_bytes(rng.bytes(48), 'little') % o.P for _ in range(32)]
a = b''.join(o.fp_to_mont_bytes(x) for x in xs)
for op in (1, 5, 6, 7):
    for _ in range(20):
        lib.dbg_fp_op(op, a, a)
    t = time.perf_counter()
    for _ in range(300):
        lib.dbg_fp_op(op, a, a)
    print('op', op, '%.1f us per call' % ((time.perf_counter() - t) / 300 * 1e6))
