"""Run a small G1 MSM (host call, plain handle) a few times, for ncu launch lists: python tools/msm_run_small.py N"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cref
from crypto_b200 import lib
n = int(sys.argv[1])
lib.init()
ks = cref.random_scalars(n, 1); ss = np.array(cref.random_scalars(n, 2))
hb = lib.Bases(cref.g1_generator_muls(ks))
for _ in range(3):
    lib.msm(hb, ss)
