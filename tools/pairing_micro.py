"""Micro-timings of the pairing engine pieces (host calls, best of 5)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cref
from crypto_b200 import lib
lib.init()
ps = cref.g1_generator_muls(cref.random_scalars(2, 41)); qs = cref.g2_generator_muls(cref.random_scalars(2, 42))
gt = lib.multi_pairing(ps, qs)
def best(fn, reps=5):
    fn(); b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t)
    return b * 1e3
sq_only = (1 << 254).to_bytes(32, 'little')
ones = ((1 << 255) - 1).to_bytes(32, 'little')
t_sq = best(lambda: lib.gt_pow(gt, sq_only)); t_all = best(lambda: lib.gt_pow(gt, ones)); t_mul = best(lambda: lib.fp12_mul(gt, gt))
print('gt_pow 255 sqr + 1 mul: %.3f ms -> %.2f us / cyc_sqr' % (t_sq, (t_sq - t_mul) / 255 * 1e3))
print('gt_pow 255 sqr + 255 mul: %.3f ms -> %.2f us / f12_mul' % (t_all, (t_all - t_sq) / 254 * 1e3))
print('fp12_mul host call: %.3f ms' % t_mul)
ml = lib.multi_miller_loop(ps, qs)
print('multi_miller_loop(2): %.3f ms   final_exponentiation: %.3f ms   multi_pairing(2): %.3f ms' % (
    best(lambda: lib.multi_miller_loop(ps, qs)), best(lambda: lib.final_exponentiation(ml)), best(lambda: lib.multi_pairing(ps, qs))))
x = cref.g1_generator_muls(cref.random_scalars(64, 1))[:48 * 64]
print('fp inverse x64 fermat %.3f ms  binary %.3f ms' % (best(lambda: lib.dbg_fp_op(5, x, x)), best(lambda: lib.dbg_fp_op(6, x, x))))
