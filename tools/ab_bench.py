"""A/B timing of one library build: python tools/ab_bench.py   (DOCKGPU_LIB=/path/to/other/libdockgpu.so selects the build)
Prints one JSON line: G1 MSM 2^20 (raw bases and resident table, device-resident operands), batch_mul and the fused
accumulator update at 10 000 elements, deserialize_compressed of 2^16 G1 points, G2 MSM 2^18 and a 2-pair product check."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cref
from crypto_b200 import lib


def dev_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def host_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t)
    return best * 1e3


lib.init()
res = {'lib': os.environ.get('DOCKGPU_LIB', 'crypto_b200/libdockgpu.so')}
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
one = np.zeros(32, np.uint8); one[0] = 1
# G1 MSM 2^20
n = 1 << 20
sc = cref.random_scalars(n, 320); ks = cref.random_scalars(n, 420)
tbl = lib.FixedBaseTable(cref.g1_generator_muls(one), n)
bases = np.array(tbl.mul_many_normalized(ks)); tbl.free()
hb = lib.Bases(bases)
d_s = torch.from_numpy(np.array(sc)).cuda(); d_o = torch.zeros(144, dtype=torch.uint8, device='cuda')
res['g1_msm_2p20_raw_ms'] = dev_ms(lambda: lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream))
hb.precompute()
res['g1_msm_2p20_table_ms'] = dev_ms(lambda: lib.msm_handle_device(hb, d_s.data_ptr(), n, d_o.data_ptr(), ts.cuda_stream))
hb.free()
# batch mul / fused update, 10 000 elements
m = 10000
pts = bases[: 96 * m].copy(); sa = np.array(sc[: 32 * m]).copy(); sb = np.array(ks[: 32 * m]).copy()
v = bases[96 * m: 96 * (m + 1)].copy()
res['batch_mul_10k_ms'] = host_ms(lambda: lib.batch_mul(pts, sa))
res['fused_update_10k_ms'] = host_ms(lambda: lib.batch_mul_add_same_g1(pts, sa, v, sb))
# fixed-base multiply_many 2^18 (CRS generator shape)
tbl = lib.FixedBaseTable(cref.g1_generator_muls(one), 1 << 18)
s18 = np.array(sc[: 32 << 18]).copy()
res['fixed_base_2p18_normalized_ms'] = host_ms(lambda: tbl.mul_many_normalized(s18), reps=3)
tbl.free()
# deserialize_compressed 2^16 G1
comp = np.array(lib.serialize_points(bases[: 96 << 16], compressed=True))
res['deserialize_2p16_ms'] = host_ms(lambda: lib.deserialize_points(comp, compressed=True, validate=True), reps=3)
# G2 MSM 2^18
n2 = 1 << 18
tbl2 = lib.FixedBaseTable(cref.g2_generator_muls(one), n2, g2=True)
b2 = np.array(tbl2.mul_many_normalized(np.array(ks[: 32 * n2]).copy())); tbl2.free()
hb2 = lib.Bases(b2, g2=True)
d_o2 = torch.zeros(288, dtype=torch.uint8, device='cuda')
res['g2_msm_2p18_raw_ms'] = dev_ms(lambda: lib.msm_handle_device(hb2, d_s.data_ptr(), n2, d_o2.data_ptr(), ts.cuda_stream))
hb2.precompute()
res['g2_msm_2p18_table_ms'] = dev_ms(lambda: lib.msm_handle_device(hb2, d_s.data_ptr(), n2, d_o2.data_ptr(), ts.cuda_stream))
hb2.free()
# 2-pair product check
g1p = bases[: 192].copy(); g2p = b2[: 384].copy()
res['pairing_check_2_ms'] = host_ms(lambda: lib.multi_pairing_is_one(g1p, g2p), reps=10)
print(json.dumps(res), flush=True)
