"""The reference issues independent MSMs concurrently (`join!` -> rayon::join, utils/src/macros.rs:68-84).  The ABI gives
every calling thread its own stream and arena, so two host threads overlap their MSMs on one GPU: python tools/join_probe.py N"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cref
from crypto_b200 import lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
lib.init()
one = np.zeros(32, np.uint8); one[0] = 1
tbl = lib.FixedBaseTable(cref.g1_generator_muls(one), n)
hs, ss = [], []
for k in range(4):
    ks = cref.random_scalars(n, 10 + k)
    hs.append(lib.Bases(np.array(tbl.mul_many_normalized(ks)))); ss.append(np.array(cref.random_scalars(n, 20 + k)))
tbl.free()
outs = [None] * 4
def work(k): outs[k] = lib.msm(hs[k], ss[k])
# persistent workers, like a rayon pool: a thread's stream, arena and pinned status word are created on its first call
import queue
qs = [queue.Queue() for _ in range(4)]; done = queue.Queue()
def loop(k):
    while True:
        j = qs[k].get()
        if j is None: return
        work(j); done.put(j)
pool = [threading.Thread(target=loop, args=(k,), daemon=True) for k in range(4)]
for t in pool: t.start()
def seq(m):
    for k in range(m): work(k)
def par(m):
    for k in range(m): qs[k].put(k)
    for k in range(m): done.get()
def best(fn, *a):
    fn(*a); fn(*a); b = 1e9
    for _ in range(10):
        t = time.perf_counter(); fn(*a); b = min(b, time.perf_counter() - t)
    return b * 1e3
for m in (1, 2, 4):
    print('%d MSMs of %d terms: one after the other %.3f ms, from %d pool threads %.3f ms' % (m, n, best(seq, m), m, best(par, m)), flush=True)
for k in range(4): qs[k].put(None)
