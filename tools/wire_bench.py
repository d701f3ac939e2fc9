"""Row f4 timing: deserialize_compressed (Validate::Yes) of 2^LOGN G1 points, GPU host call vs the C restatement on
the host cores; G2 GPU-only.  python tools/wire_bench.py [LOGN] [--out FILE]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cref
from crypto_b200 import lib

logn = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 18
n = 1 << logn
lib.init()
res = {'points': n, 'host_cores': os.cpu_count()}
aff = cref.g1_generator_muls(cref.random_scalars(n, 91))


def best(fn, reps=3):
    fn()
    b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); b = min(b, time.perf_counter() - t)
    return b, r


t_ser, enc = best(lambda: lib.serialize_points(aff))
t_de, (out, st, bad) = best(lambda: lib.deserialize_points(enc, validate=True))
t_de_nv, _ = best(lambda: lib.deserialize_points(enc, validate=False))
assert bad == 0 and np.array_equal(out, aff)
t0 = time.perf_counter(); cout, cst = cref.g1_deserialize_compressed(enc, True); t_cpu = time.perf_counter() - t0
t0 = time.perf_counter(); cref.g1_deserialize_compressed(enc, False); t_cpu_nv = time.perf_counter() - t0
assert np.array_equal(cout, aff) and not cst.any()
res['g1'] = {'serialize_compressed_gpu_ms': t_ser * 1e3, 'deserialize_compressed_validated_gpu_ms': t_de * 1e3,
             'deserialize_compressed_unvalidated_gpu_ms': t_de_nv * 1e3, 'deserialize_compressed_validated_cpu_ms': t_cpu * 1e3,
             'deserialize_compressed_unvalidated_cpu_ms': t_cpu_nv * 1e3, 'ok': True}
m = n >> 2
aff2 = cref.g2_generator_muls(cref.random_scalars(m, 92))
t_ser, enc2 = best(lambda: lib.serialize_points(aff2, g2=True))
t_de, (out2, st2, bad2) = best(lambda: lib.deserialize_points(enc2, g2=True, validate=True))
assert bad2 == 0 and np.array_equal(out2, aff2)
t0 = time.perf_counter(); cout2, cst2 = cref.g2_deserialize_compressed(enc2, True); t_cpu2 = time.perf_counter() - t0
assert np.array_equal(cout2, aff2) and not cst2.any()
res['g2'] = {'points': m, 'serialize_compressed_gpu_ms': t_ser * 1e3, 'deserialize_compressed_validated_gpu_ms': t_de * 1e3,
             'deserialize_compressed_validated_cpu_ms': t_cpu2 * 1e3, 'ok': True}
res['note'] = 'gpu_ms = host C-ABI call incl. H2D/D2H; cpu_ms = oracle C restatement (OpenMP over points) on the host cores'
s = json.dumps(res, indent=1)
if '--out' in sys.argv:
    open(sys.argv[sys.argv.index('--out') + 1], 'w').write(s)
print(s)
