"""Run the device-chained LegoGroth16 prover call (dg_groth16_prove_msms) a few times, for ncu launch lists and timing:
python tools/prover_run.py LOGD [table|plain] [streams]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crypto_b200 import lib, groth16 as g16, group as gp
from oracle import bls12_381 as o
from tools.synth_circuit import synthetic_r1cs
logd = int(sys.argv[1]); table = len(sys.argv) > 2 and sys.argv[2] == 'table'
lib.init()
cs, w = synthetic_r1cs((1 << logd) - 3, num_public=2, seed=3)
g1, g2 = o.g1_to_bytes(o.G1_GEN), o.g2_to_bytes(o.G2_GEN)
pk, ni = g16.generate_parameters(cs, 0x1111, 0x2222, 0x3333, 0x4444, 0x5555, 0x1234567, g1, g2, 2)
w_mont = gp.fr_to_mont(w)
dpk = g16.DeviceProvingKey(pk, cs, precompute=table)
nw, cw = cs.num_witness_variables, pk.vk.commit_witness_count
jobs = [(dpk.l_query, ni + cw, nw - cw), (dpk.a_query, 0, ni + nw), (dpk.b_g1_query, 0, ni + nw), (dpk.b_g2_query, 0, ni + nw),
        (dpk.gamma_abc_committed, ni, cw)]
print('MARK warmup', flush=True)
if len(sys.argv) > 3:
    lib.dbg_set_tunable(1, int(sys.argv[3]))          # 2: two streams
for _ in range(2):
    lib.groth16_prove_msms(dpk.r1cs, w_mont, dpk.h_query, jobs)
ts = []
for _ in range(6):
    t = time.perf_counter(); lib.groth16_prove_msms(dpk.r1cs, w_mont, dpk.h_query, jobs); ts.append(time.perf_counter() - t)
print('chained call ms:', [round(1e3 * x, 3) for x in ts], flush=True)
