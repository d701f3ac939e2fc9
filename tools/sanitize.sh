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
print('sanitizer workload ok =', bool(ok))
PY
for tool in memcheck racecheck; do
  $SAN --tool $tool --error-exitcode 1 python /tmp/san_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
