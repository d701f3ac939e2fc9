#!/bin/bash
# compute-sanitizer memcheck + racecheck over the window-group split of the MSM only (the full workload is tools/sanitize.sh)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/san_split.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import cref
from crypto_b200 import lib
lib.init(0)
n = 600
ks = cref.random_scalars(n, 1); ss = cref.random_scalars(n, 2)
bases = cref.g1_generator_muls(ks)
b2 = cref.g2_generator_muls(ks[:32 * 40])
exp = bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
ok = True
for h in (2, 3, 7):
    lib.dbg_set_tunable(2, h)
    ok &= bytes(cref.normalize_batch_g1(lib.msm(bases, ss))) == exp
    hb = lib.Bases(bases)
    ok &= bytes(cref.normalize_batch_g1(lib.msm(hb, ss))) == exp
    lib.msm_set_affine_rounds(2)
    ok &= bytes(cref.normalize_batch_g1(lib.msm(hb, ss))) == exp
    lib.msm_set_affine_rounds(-1)
    hb.free()
ok &= bytes(cref.normalize_batch_g2(lib.msm(b2, ss[:32 * 40], g2=True))) == bytes(cref.normalize_batch_g2(cref.msm_g2(b2, ss[:32 * 40])))
lib.dbg_set_tunable(2, 0)
print('sanitizer split workload ok =', bool(ok))
PY
for tool in memcheck racecheck; do
  $SAN --tool $tool --error-exitcode 1 python /tmp/san_split.py > gpurun_out/sanitizer_split_$tool.log 2>&1
  echo "$tool exit=$?"; grep -c "k_window_combine\|k_digits" gpurun_out/sanitizer_split_$tool.log; tail -3 gpurun_out/sanitizer_split_$tool.log
done
