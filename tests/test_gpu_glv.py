"""GPU: the GLV split of the plain-bases MSM path (csrc/msm_kernels.cuh glv_split / k_glv_expand): the same group
element as the unsplit path, the oracle and the known-discrete-log identity, for G1 and G2, raw bases and resident plain
handles, including the scalars where the split has its corner cases (multiples of x^2, x^2 / 2, lambda, r - 1 ...)."""
import numpy as np
import pytest

from crypto_b200 import msm
from tests import helpers as h

pytestmark = pytest.mark.gpu

R = msm.R_MODULUS
EDGE = [0, 1, 2, R - 1, R - 2, (R - 1) // 2, (R + 1) // 2, msm.GLV_X2, msm.GLV_X2 - 1, msm.GLV_X2 + 1, msm.GLV_X2 // 2,
        msm.GLV_X2 // 2 + 1, msm.GLV_LAMBDA, msm.GLV_LAMBDA + 1, R - msm.GLV_LAMBDA, 5 * msm.GLV_X2, 5 * msm.GLV_X2 + msm.GLV_X2 // 2 + 1,
        (1 << 127), (1 << 127) - 1, (1 << 128) + 1, (1 << 254) - 1, 1 << 254]


def _glv(dg, on):
    dg.dbg_set_tunable(4, 0 if on else 1)


@pytest.mark.parametrize('n', [1, 31, 33, 1000, 1 << 14, (1 << 17) + 5])
def test_g1_glv_equals_unsplit_and_known_dlog(dg, cref, n):
    bases, ks = h.g1_bases(n, 600 + n)
    ss = np.array(h.rand_scalars(n, 601 + n))
    edge = h.scalars_bytes([e % R for e in EDGE])[:32 * min(n, len(EDGE))]
    ss[:len(edge)] = edge
    exp = h.known_dlog_msm_g1(ks, ss)
    try:
        _glv(dg, True)
        assert dg.msm_plan(n)[0] >= 4
        got_glv = h.affine_g1(dg.msm(bases, ss))
        hb = dg.Bases(bases)
        got_handle = h.affine_g1(dg.msm(hb, ss))
        got_prefix = h.affine_g1(dg.msm(hb, ss[:32 * (n // 2)])) if n > 1 else None
        hb.free()
        _glv(dg, False)
        got_plain = h.affine_g1(dg.msm(bases, ss))
    finally:
        _glv(dg, True)
    assert got_glv == got_plain == got_handle == exp
    if n > 1:
        assert got_prefix == h.known_dlog_msm_g1(ks[:32 * (n // 2)], ss[:32 * (n // 2)])
    if n <= 1000:
        assert got_glv == h.affine_g1(cref.msm_g1(bases, ss))


@pytest.mark.parametrize('n', [1, 40, 3001])
def test_g2_glv_equals_unsplit_and_known_dlog(dg, cref, n):
    bases, ks = h.g2_bases(n, 700 + n)
    ss = np.array(h.rand_scalars(n, 701 + n))
    edge = h.scalars_bytes([e % R for e in EDGE])[:32 * min(n, len(EDGE))]
    ss[:len(edge)] = edge
    exp = h.known_dlog_msm_g2(ks, ss)
    try:
        _glv(dg, True)
        got_glv = h.affine_g2(dg.msm(bases, ss, g2=True))
        hb = dg.Bases(bases, g2=True)
        got_handle = h.affine_g2(dg.msm(hb, ss, g2=True))
        hb.free()
        _glv(dg, False)
        got_plain = h.affine_g2(dg.msm(bases, ss, g2=True))
    finally:
        _glv(dg, True)
    assert got_glv == got_plain == got_handle == exp
    # the G2 line sums of the bucket reduction on 12-lane quads (A/B switch; the default keeps 4-lane quads there) and
    # every forced window from 4 to 16 bits (10 / 13 / 16 are the automatic choices) give the same point
    try:
        dg.dbg_set_tunable(6, 2)
        assert h.affine_g2(dg.msm(bases, ss, g2=True)) == exp
        dg.dbg_set_tunable(6, 0)
        for c in (4, 9, 13, 16):
            dg.msm_set_window(c)
            assert h.affine_g2(dg.msm(bases, ss, g2=True)) == exp
    finally:
        dg.dbg_set_tunable(6, 0)
        dg.msm_set_window(0)


def test_glv_with_identity_equal_and_opposite_bases(dg, cref):
    """phi maps the identity record to itself and commutes with negation: the adversarial base sets of the plain path."""
    n = 64
    bases, ks = h.g1_bases(n, 55)
    b = bytearray(bytes(bases))
    b[0:96] = bytes(96)                                  # identity
    b[96 * 5:96 * 6] = b[96 * 4:96 * 5]                  # P, P
    b[96 * 7:96 * 8] = h.neg_g1(bytes(b[96 * 6:96 * 7]))  # P, -P
    ss = h.rand_scalars(n, 56)
    assert h.affine_g1(dg.msm(bytes(b), ss)) == h.affine_g1(cref.msm_g1(np.frombuffer(bytes(b), dtype=np.uint8), ss))
    same = np.frombuffer(bytes(b[96:192]) * n, dtype=np.uint8)       # every base equal: one hot bucket per window
    assert h.affine_g1(dg.msm(same, ss)) == h.affine_g1(cref.msm_g1(same, ss))


@pytest.mark.parametrize('g2', [False, True])
def test_batch_mul_glv_edge_scalars_and_non_field_bigints(dg, cref, g2):
    """dg_batch_mul: canonical scalars take the GLV path, integers >= r (legal BigInts for mul_bigint) the generic
    64-digit path; both equal the oracle's double-and-add, identity points included."""
    big = [R, R + 1, (1 << 255) - 1, (1 << 255) - 19, R + msm.GLV_X2, 2 * R - 1]
    scal = [e % (1 << 255) for e in EDGE] + big + h.ints_of(h.rand_scalars(40, 91))
    n = len(scal)
    pts, _ = (h.g2_bases if g2 else h.g1_bases)(n, 92)
    rec = 192 if g2 else 96
    b = bytearray(bytes(pts))
    b[rec * 3:rec * 4] = bytes(rec)                       # one identity point
    sb = np.frombuffer(b''.join(int(s).to_bytes(32, 'little') for s in scal), dtype=np.uint8)
    exp = (cref.batch_mul_g2 if g2 else cref.batch_mul_g1)(np.frombuffer(bytes(b), dtype=np.uint8), sb)
    aff = h.affine_g2 if g2 else h.affine_g1
    try:
        _glv(dg, True)
        got = dg.batch_mul(bytes(b), sb, g2=g2)
        _glv(dg, False)
        got_plain = dg.batch_mul(bytes(b), sb, g2=g2)
    finally:
        _glv(dg, True)
    assert aff(got) == aff(exp) == aff(got_plain)


@pytest.mark.parametrize('n,hwin', [(33, 2), (1000, 2), (1000, 7), (1 << 14, 3), ((1 << 17) + 5, 4)])
def test_window_group_split_equals_single_stream(dg, cref, n, hwin):
    """tunable 2 (csrc/msm_host.cuh msm_run): the high windows on a second stream, the low group adds their sum.  Same group
    element as the single-stream path, for raw bases, a resident handle, forced batch-affine rounds and a scalar >= r
    (rejected by either group's digit pass)."""
    bases, ks = h.g1_bases(n, 900 + n)
    ss = np.array(h.rand_scalars(n, 901 + n))
    edge = h.scalars_bytes([e % R for e in EDGE])[:32 * min(n, len(EDGE))]
    ss[:len(edge)] = edge
    exp = h.known_dlog_msm_g1(ks, ss)
    bad = np.array(ss)
    bad[32 * (n - 1):32 * n] = 0xff
    try:
        dg.dbg_set_tunable(2, hwin)
        got_raw = h.affine_g1(dg.msm(bases, ss))
        hb = dg.Bases(bases)
        got_handle = h.affine_g1(dg.msm(hb, ss))
        dg.msm_set_affine_rounds(2)
        got_rounds = h.affine_g1(dg.msm(hb, ss))
        dg.msm_set_affine_rounds(-1)
        with pytest.raises(dg.DockGpuError):
            dg.msm(hb, bad)
        got_after = h.affine_g1(dg.msm(hb, ss))          # the error word of the failed call does not leak into the next
        hb.free()
    finally:
        dg.msm_set_affine_rounds(-1)
        dg.dbg_set_tunable(2, 0)
    assert got_raw == got_handle == got_rounds == got_after == exp
    if n <= 1000:
        assert got_raw == h.affine_g1(cref.msm_g1(bases, ss))


def test_window_group_split_g2(dg, cref):
    n = 3001
    bases, ks = h.g2_bases(n, 950)
    ss = np.array(h.rand_scalars(n, 951))
    exp = h.known_dlog_msm_g2(ks, ss)
    try:
        dg.dbg_set_tunable(2, 3)
        got = h.affine_g2(dg.msm(bases, ss, g2=True))
        dg.msm_set_affine_rounds(2)
        got_rounds = h.affine_g2(dg.msm(bases, ss, g2=True))
    finally:
        dg.msm_set_affine_rounds(-1)
        dg.dbg_set_tunable(2, 0)
    assert got == got_rounds == exp
