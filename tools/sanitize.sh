#!/bin/bash
# compute-sanitizer passes over small configurations (race detection / sanitizers).
# usage: tools/sanitize.sh  (on a GPU box); logs under gpurun_out/
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/san_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import cref
from crypto_b200 import lib
lib.init(0)
n = 600
ks = cref.random_scalars(n, 1); ss = cref.random_scalars(n, 2)
bases = cref.g1_generator_muls(ks)
ok = bytes(cref.normalize_batch_g1(lib.msm(bases, ss))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
lib.msm_set_affine_rounds(3)          # small inputs do not reach the batch-affine stage on their own: force it
ok &= bytes(cref.normalize_batch_g1(lib.msm(bases, ss))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
eq = np.tile(ss[:32], n)              # all-equal scalars: one hot bucket per window, doubling / cancellation paths stay cold
ok &= bytes(cref.normalize_batch_g1(lib.msm(bases, eq))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, eq)))
lib.msm_set_affine_rounds(0)          # ... and without the rounds the hot bucket spans > 24 chunks: k_fixup_long_part / _final
ok &= bytes(cref.normalize_batch_g1(lib.msm(bases, eq))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, eq)))
hb = lib.Bases(bases).precompute(10)
ok &= bytes(cref.normalize_batch_g1(lib.msm(hb, ss))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
hb.free()
b2 = cref.g2_generator_muls(ks[:32 * 40])
ok &= bytes(cref.normalize_batch_g2(lib.msm(b2, ss[:32 * 40], g2=True))) == bytes(cref.normalize_batch_g2(cref.msm_g2(b2, ss[:32 * 40])))
lib.msm_set_affine_rounds(-1)
enc = lib.serialize_points(bases[:96 * 64])
dec, st, bad = lib.deserialize_points(enc, validate=True)
ok &= bad == 0 and bytes(dec) == bytes(bases[:96 * 64])
enc2 = lib.serialize_points(b2[:192 * 8], g2=True, compressed=False)
dec2, st2, bad2 = lib.deserialize_points(enc2, g2=True, compressed=False, validate=True)
ok &= bad2 == 0 and bytes(dec2) == bytes(b2[:192 * 8])
t = lib.FixedBaseTable(bases[:96], 40)
ok &= bytes(cref.normalize_batch_g1(t.mul_many(ss[:32 * 8]))) == bytes(cref.normalize_batch_g1(cref.batch_mul_g1(np.tile(bases[:96], 8), ss[:32 * 8])))
t.free()
ok &= bytes(lib.multi_pairing(bases[:96 * 2], b2[:192 * 2])) == bytes(cref.multi_pairing(bases[:96 * 2], b2[:192 * 2]))
d = cref.random_scalars(1 << 10, 5)
ok &= bytes(lib.fr_ntt(d, 10, False, True)) == bytes(cref.fr_ntt(d, 10, False, True))
# ---- round 2: GLV on / off, sharded entry point, windowed batch multiplication (GLV and generic chain), fused update
# (joint chain and table), Pornin inversion, compress, chained Groth16 prover on two streams
lib.dbg_set_tunable(4, 1)
ok &= bytes(cref.normalize_batch_g1(lib.msm(bases, ss))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
lib.dbg_set_tunable(4, 0)
lib.dbg_set_tunable(2, 3)             # window-group split: two scratch regions, two streams, joined in k_window_combine
ok &= bytes(cref.normalize_batch_g1(lib.msm(bases, ss))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
lib.msm_set_affine_rounds(2)
ok &= bytes(cref.normalize_batch_g1(lib.msm(bases, ss))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
ok &= bytes(cref.normalize_batch_g2(lib.msm(b2, ss[:32 * 40], g2=True))) == bytes(cref.normalize_batch_g2(cref.msm_g2(b2, ss[:32 * 40])))
lib.msm_set_affine_rounds(-1)
lib.dbg_set_tunable(2, 0)
ok &= bytes(cref.normalize_batch_g1(lib.msm_sharded(bases, ss))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
hs = lib.ShardedBases(bases); hs.precompute(10)
ok &= bytes(cref.normalize_batch_g1(lib.msm_sharded(hs, ss))) == bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
hs.free()
big = np.array(ss[:32 * 40]); big[:32] = 0xff; big[31] = 0x7f          # one integer >= r: generic 65-digit chain
ok &= bytes(cref.normalize_batch_g1(lib.batch_mul(bases[:96 * 40], big))) == bytes(cref.normalize_batch_g1(cref.batch_mul_g1(bases[:96 * 40], big)))
ok &= bytes(cref.normalize_batch_g2(lib.batch_mul(b2, ss[:32 * 40], g2=True))) == bytes(cref.normalize_batch_g2(cref.batch_mul_g2(b2, ss[:32 * 40])))
tv = lib.FixedBaseTable(bases[:96], 40)
lib.dbg_set_tunable(5, 2)
u_tab = bytes(lib.batch_mul_add_fixed_g1(bases[:96 * 40], ss[:32 * 40], tv, ss[32 * 40:32 * 80]))
lib.dbg_set_tunable(5, 0)
tv.free()
ok &= u_tab == bytes(lib.batch_mul_add_same_g1(bases[:96 * 40], ss[:32 * 40], bases[:96], ss[32 * 40:32 * 80]))
x48 = bases[:48 * 64]
ok &= bytes(lib.dbg_fp_op(3, x48, x48)) == bytes(lib.dbg_fp_op(0, x48, x48))      # dedicated squaring == multiplier
ok &= bytes(lib.dbg_fp_op(7, x48, x48)) == bytes(lib.dbg_fp_op(5, x48, x48))
ok &= len(bytes(lib.compress(bases[:96 * 20], bases[96 * 20:96 * 40], ss[:32]))) == 96 * 20
sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
from crypto_b200 import groth16 as g16
from oracle import bls12_381 as o
from tools.synth_circuit import synthetic_r1cs
cs, w = synthetic_r1cs(200, seed=4)
pk, ni = g16.generate_parameters(cs, 11, 12, 13, 14, 15, 99991, o.g1_to_bytes(o.G1_GEN), o.g2_to_bytes(o.G2_GEN), 2)
dpk = g16.DeviceProvingKey(pk, cs)
proof, _ = g16.create_proof(dpk, w, 5, 6, 7)
ok &= g16.verify_proof(g16.prepare_verifying_key(pk.vk), proof, w[1:ni])
dpk.free()
print('sanitizer workload ok =', bool(ok))
PY
for tool in memcheck racecheck synccheck; do
  EXTRA=""
  # synccheck cannot track the one-mbarrier-per-thread staging of k_accumulate (it overflows its barrier table and the
  # launch then fails inside the tool): that kernel is checked by memcheck / racecheck only
  if [ $tool = synccheck ]; then EXTRA="--kernel-name-exclude kns=k_accumulate"; fi
  $SAN --tool $tool $EXTRA --error-exitcode 1 python /tmp/san_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
