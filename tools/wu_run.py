"""Run the fused accumulator witness update (10 000 elements) and batch_mul a few times, for ncu launch lists:
python tools/wu_run.py"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import cref
from crypto_b200 import lib
lib.init()
m = 10000
ks = cref.random_scalars(m + 1, 1); sa = np.array(cref.random_scalars(m, 2)); sb = np.array(cref.random_scalars(m, 3))
one = np.zeros(32, np.uint8); one[0] = 1
tbl = lib.FixedBaseTable(cref.g1_generator_muls(one), m + 1)
pts = np.array(tbl.mul_many_normalized(ks)); tbl.free()
v = pts[96 * m:].copy(); pts = pts[:96 * m].copy()
for _ in range(3):
    lib.batch_mul_add_same_g1(pts, sa, v, sb)
for _ in range(2):
    lib.batch_mul(pts, sa)
