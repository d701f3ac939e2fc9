"""Times the device-resident Fr NTT and the witness-map tail next to the CPU oracle.
python tools/ntt_bench.py [--out FILE]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib

lib.init()
out = {'host_cores': len(os.sched_getaffinity(0)), 'ntt': [], 'note': 'gpu_ms: device-resident data, CUDA events, mean of 20 after warm-up; '
       'algorithmic bytes = 2 * n * 32 B (read + write once); cpu_ms: oracle C NTT (OpenMP over butterflies)'}
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for logn in (16, 18, 19, 20, 22):
    n = 1 << logn
    data = cref.random_scalars(n, logn)
    d = torch.from_numpy(data).cuda(); tmp = torch.empty_like(d)
    for _ in range(3):
        lib.fr_ntt_device(d.data_ptr(), tmp.data_ptr(), logn, False, False, ts.cuda_stream)
    torch.cuda.synchronize()
    res = {}
    for name, inv, cos in (('fft', False, False), ('coset_ifft', True, True)):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            lib.fr_ntt_device(d.data_ptr(), tmp.data_ptr(), logn, inv, cos, ts.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        res[name + '_gpu_ms'] = e0.elapsed_time(e1) / 20
    t = time.perf_counter(); cref.fr_ntt(data, logn); cpu = time.perf_counter() - t
    ent = {'logn': logn, **res, 'cpu_ms': cpu * 1e3, 'alg_GBps_fft': 2 * n * 32 / (res['fft_gpu_ms'] * 1e-3) / 1e9}
    out['ntt'].append(ent)
    print(ent, flush=True)
# witness-map tail at 2^18 through the host call (H2D of a, b, c and D2H of h included)
logn = 18; n = 1 << logn
a, b, c = (cref.random_scalars(n, s) for s in (1, 2, 3))
lib.qap_h_from_abc(a, b, c, logn)
t = time.perf_counter(); h = lib.qap_h_from_abc(a, b, c, logn); g = time.perf_counter() - t
t = time.perf_counter(); hc = cref.qap_h_from_abc(a, b, c, logn); cpu = time.perf_counter() - t
out['qap_h_from_abc_2^18'] = {'gpu_ms_host_call': g * 1e3, 'cpu_ms': cpu * 1e3, 'ok': bool(np.array_equal(h, hc))}
print(out['qap_h_from_abc_2^18'])
if len(sys.argv) > 2 and sys.argv[1] == '--out':
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
