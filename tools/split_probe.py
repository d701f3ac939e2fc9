"""Does splitting one MSM by base range over concurrent streams (host threads, each with its own stream and arena) pay on
ONE GPU?  python tools/split_probe.py LOGN PARTS"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cref
from crypto_b200 import lib
logn = int(sys.argv[1]); parts = int(sys.argv[2])
n = 1 << logn
lib.init()
sc = np.array(cref.random_scalars(n, 320)); ks = cref.random_scalars(n, 420)
one = np.zeros(32, np.uint8); one[0] = 1
tbl = lib.FixedBaseTable(cref.g1_generator_muls(one), n)
bases = np.array(tbl.mul_many_normalized(ks)); tbl.free()
import torch
pin = torch.from_numpy(sc).pin_memory().numpy()
whole = lib.Bases(bases)
def run_whole():
    return lib.msm(whole, pin)
m = n // parts
hs = [lib.Bases(bases[96 * m * i: 96 * m * (i + 1)]) for i in range(parts)]
outs = [None] * parts
def worker(i):
    outs[i] = lib.msm(hs[i], pin[32 * m * i: 32 * m * (i + 1)])
def run_split():
    th = [threading.Thread(target=worker, args=(i,)) for i in range(parts)]
    for t in th: t.start()
    for t in th: t.join()
    return lib.fold(np.concatenate([np.frombuffer(bytes(o), np.uint8) for o in outs]))
def best(fn, reps=8):
    for _ in range(3): fn()
    b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t)
    return b * 1e3
a = bytes(lib.normalize_batch(run_whole())); b = bytes(lib.normalize_batch(run_split()))
print('equal', a == b, 'whole %.3f ms   split x%d %.3f ms' % (best(run_whole), parts, best(run_split)))
