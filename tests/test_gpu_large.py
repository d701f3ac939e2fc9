"""GPU: the MSM at the full BASELINE sizes that the oracle's own Pippenger cannot finish in seconds -- G1 at 2^22 and
2^24 terms (configs 2 and 5), G2 at 2^18 (the slowest MSM of config 3) -- checked bit-exact through the O(n)
known-discrete-log identity  sum s_i (k_i G) = (sum s_i k_i mod r) G.  The bases k_i G are produced on the GPU by the
fixed-base path (itself parity-tested against the oracle in test_gpu_parity.py / test_gpu_checkers.py); the scalar side
of the identity is the oracle's 576-bit dot product.  2^24 runs the 22 GB-scratch plan nothing smaller reaches."""
import numpy as np
import pytest

from oracle import bls12_381 as o
from tests import helpers as h

pytestmark = pytest.mark.gpu


def _gpu_bases(dg, ks, g2=False):
    one = np.zeros(32, np.uint8)
    one[0] = 1
    gen = (h.cref.g2_generator_muls if g2 else h.cref.g1_generator_muls)(one)
    tbl = dg.FixedBaseTable(gen, max(len(ks) // 32, 32), g2=g2)
    out = np.array(tbl.mul_many_normalized(ks))
    tbl.free()
    return out


def _expected(ks, ss, g2=False):
    tot = h.cref.scalar_dot_mod_r(ks, ss)
    return bytes((h.cref.g2_generator_muls if g2 else h.cref.g1_generator_muls)(h.scalars_bytes([tot])))


@pytest.mark.parametrize('logn', [22, 24])
def test_g1_msm_full_size_raw_and_resident_table(dg, cref, logn):
    n = 1 << logn
    ks, ss = h.rand_scalars(n, 7000 + logn), h.rand_scalars(n, 7100 + logn)
    bases = _gpu_bases(dg, ks)
    # spot-check the generated bases against the oracle
    assert bytes(bases[:96 * 4]) == bytes(cref.g1_generator_muls(ks[:32 * 4]))
    assert bytes(bases[-96:]) == bytes(cref.g1_generator_muls(ks[-32:]))
    exp = _expected(ks, ss)
    assert h.affine_g1(dg.msm(bases, ss)) == exp                       # raw bases from host memory (GLV path)
    hb = dg.Bases(bases)
    assert h.affine_g1(dg.msm(hb, ss)) == exp                          # resident plain handle
    hb.precompute(0)
    c, rounds = dg.msm_plan(n, precomputed_c=20)
    assert c == 20 and rounds >= 1
    assert h.affine_g1(dg.msm(hb, ss)) == exp                          # 13-row table, c = 20
    m = n - 12345                                                       # a prefix that is not a power of two
    assert h.affine_g1(dg.msm(hb, ss[:32 * m])) == _expected(ks[:32 * m], ss[:32 * m])
    hb.free()


def test_g2_msm_2p18_raw_and_resident_table(dg, cref):
    n = 1 << 18
    ks, ss = h.rand_scalars(n, 7200), np.array(h.rand_scalars(n, 7201))
    # witness-shaped scalars (config 3): 10 % zeros, 10 % ones, 10 % 16-bit values
    sv = ss.reshape(n, 32)
    sv[: n // 10] = 0
    sv[n // 10: n // 5] = 0
    sv[n // 10: n // 5, 0] = 1
    sv[n // 5: 3 * n // 10, 2:] = 0
    bases = _gpu_bases(dg, ks, g2=True)
    assert bytes(bases[:192 * 2]) == bytes(cref.g2_generator_muls(ks[:64]))
    exp = _expected(ks, ss, g2=True)
    assert h.affine_g2(dg.msm(bases, ss, g2=True)) == exp
    hb = dg.Bases(bases, g2=True)
    assert h.affine_g2(dg.msm(hb, ss, g2=True)) == exp
    hb.precompute(0)
    assert h.affine_g2(dg.msm(hb, ss, g2=True)) == exp
    hb.free()
