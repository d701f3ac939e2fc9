"""GPU parity of the batch-affine pre-reduction stage (csrc/msm_affine.cuh): every round count gives
the same group element as the oracle, including the operand pairs that are not a plain chord
addition (identity bases, P + P, P + (-P)) and scalar distributions with one hot bucket."""
import numpy as np
import pytest

from oracle import bls12_381 as o
from tests import helpers as h

pytestmark = pytest.mark.gpu


def _with_rounds(dg, r, fn):
    dg.msm_set_affine_rounds(r)
    try:
        return fn()
    finally:
        dg.msm_set_affine_rounds(-1)


@pytest.mark.parametrize('rounds', [0, 1, 2, 3, 6])
@pytest.mark.parametrize('n', [1, 2, 3, 33, 1000, 20011])
def test_rounds_g1_vs_known_dlog(dg, rounds, n):
    bases, ks = h.g1_bases(n, 5000 + n)
    ss = h.rand_scalars(n, 6000 + n)
    res = _with_rounds(dg, rounds, lambda: dg.msm(bases, ss))
    assert h.affine_g1(res) == h.known_dlog_msm_g1(ks, ss)


@pytest.mark.parametrize('rounds', [0, 1, 3, 5])
def test_rounds_g2_vs_known_dlog(dg, rounds):
    n = 3001
    bases, ks = h.g2_bases(n, 7001)
    ss = h.rand_scalars(n, 7002)
    res = _with_rounds(dg, rounds, lambda: dg.msm(bases, ss, g2=True))
    assert h.affine_g2(res) == h.known_dlog_msm_g2(ks, ss)


@pytest.mark.parametrize('rounds', [1, 2, 4, 6])
def test_rounds_adversarial_g1(dg, cref, rounds):
    n = 4096
    bases, ks = h.g1_bases(n, 15)
    bases = bases.copy()
    ss = h.rand_scalars(n, 16).copy()
    sv = ss.reshape(n, 32)
    bv = bases.reshape(n, 96)
    bv[10] = 0; bv[11] = 0; bv[12] = 0                      # identity bases ...
    sv[11] = sv[10]; sv[12] = sv[10]                        # ... meeting each other in every bucket
    sv[20] = 0; sv[21] = 0; sv[21, 0] = 1
    sv[22] = np.frombuffer((o.R - 1).to_bytes(32, 'little'), np.uint8)
    bv[100:228] = bv[100]                                   # 128 equal bases with equal scalars: doubling chains
    sv[100:228] = sv[100]
    bv[301] = np.frombuffer(h.neg_g1(bv[300]), np.uint8)    # P, -P adjacent in every bucket
    sv[301] = sv[300]
    bv[303] = np.frombuffer(h.neg_g1(bv[302]), np.uint8)    # P, -P, P
    bv[304] = bv[302]
    sv[303] = sv[302]; sv[304] = sv[302]
    sv[400:1400] = 0; sv[400:1400, 0] = 1                   # 1000 scalars equal to one: a hot bucket
    sv[1400:1700, 2:] = 0                                   # 16-bit scalars
    res = _with_rounds(dg, rounds, lambda: dg.msm(bases, ss))
    assert h.affine_g1(res) == h.affine_g1(cref.msm_g1(bases, ss))


@pytest.mark.parametrize('rounds', [2, 6])
def test_rounds_adversarial_g2(dg, cref, rounds):
    n = 600
    bases, ks = h.g2_bases(n, 25)
    bases = bases.copy()
    ss = h.rand_scalars(n, 26).copy()
    sv = ss.reshape(n, 32)
    bv = bases.reshape(n, 192)
    bv[5] = 0
    bv[50:90] = bv[50]
    sv[50:90] = sv[50]
    sv[100:400] = 0; sv[100:400, 0] = 3
    res = _with_rounds(dg, rounds, lambda: dg.msm(bases, ss, g2=True))
    assert h.affine_g2(res) == h.affine_g2(cref.msm_g2(bases, ss))


@pytest.mark.parametrize('rounds', [3, 6])
def test_rounds_everything_cancels(dg, rounds):
    n = 512
    bases, _ = h.g1_bases(n, 35)
    bv = bases.reshape(n, 96).copy()
    for i in range(0, n, 2):
        bv[i + 1] = np.frombuffer(h.neg_g1(bv[i]), np.uint8)
    ss = np.repeat(h.rand_scalars(n // 2, 36).reshape(n // 2, 32), 2, axis=0).reshape(-1)
    res = _with_rounds(dg, rounds, lambda: dg.msm(bv.reshape(-1), ss))
    assert not any(h.affine_g1(res))


def test_rounds_all_scalars_equal_precomputed(dg):
    """One hot bucket of n entries through the precomputed-table path: n/2 independent additions per round."""
    n = 1 << 14
    bases, ks = h.g1_bases(n, 45)
    ss = np.tile(h.rand_scalars(1, 46), n)
    hb = dg.Bases(bases)
    try:
        hb.precompute(16)
        for rounds in (-1, 0, 4):
            res = _with_rounds(dg, rounds, lambda: dg.msm(hb, ss))
            assert h.affine_g1(res) == h.known_dlog_msm_g1(ks, ss)
    finally:
        hb.free()


def test_plan_reports_rounds(dg):
    c, r = dg.msm_plan(1 << 20, precomputed_c=20)
    assert c == 20 and r >= 1
    c, r = dg.msm_plan(1 << 20)
    assert c == 16 and r >= 1


@pytest.mark.parametrize('c', [5, 15, 17])
def test_windows_dividing_255_bits(dg, c):
    """Windows that divide 255 put a full-width digit on top (r >> (255 - c) is ~0.9 * 2^c, above half): its carry
    must land in the extra sparse digit; the largest canonical scalars must work, inputs >= 2^255 must be rejected."""
    n = 700
    bases, ks = h.g1_bases(n, 61)
    ss = h.rand_scalars(n, 62).copy()
    sv = ss.reshape(n, 32)
    for i, v in enumerate((o.R - 1, o.R - 2, (1 << 254) + 12345, (1 << 254) - 1, o.R >> 1)):
        sv[i] = np.frombuffer(int(v).to_bytes(32, 'little'), np.uint8)
    dg.msm_set_window(c)
    try:
        assert h.affine_g1(dg.msm(bases, ss)) == h.known_dlog_msm_g1(ks, ss)
        for bad in ((1 << 255), (1 << 255) - 1):
            t = ss.copy()
            t[32 * 9:32 * 10] = np.frombuffer(int(bad).to_bytes(32, 'little'), np.uint8)
            with pytest.raises(dg.DockGpuError):
                dg.msm(bases, t)
        assert h.affine_g1(dg.msm(bases, ss)) == h.known_dlog_msm_g1(ks, ss)       # still usable afterwards
    finally:
        dg.msm_set_window(0)
    hb = dg.Bases(bases)
    try:
        hb.precompute(17)
        assert h.affine_g1(dg.msm(hb, ss)) == h.known_dlog_msm_g1(ks, ss)
    finally:
        hb.free()
