"""GPU: the reference's checker types re-expressed over the backend
(utils/src/randomized_pairing_check.rs:234-421, utils/src/randomized_mult_checker.rs:136-384,
utils/src/msm.rs:116-308) plus the BBS+-shaped workload of BASELINE config 4."""
import numpy as np
import pytest

from crypto_b200 import msm, pairing_check as pc
from oracle import bls12_381 as o
from tests import helpers as h

pytestmark = pytest.mark.gpu


def test_window_table_matches_mul_bigint(dg, cref):
    """utils/src/msm.rs timing_ark_ops: table.multiply(e) == g.mul_bigint(e)."""
    g, _ = h.g1_bases(1, 1)
    elems = h.ints_of(h.rand_scalars(30, 2))
    table = msm.WindowTable.new(len(elems), g)
    assert table.scalar_size == 255 and table.window_size == 3 and table.num_windows == 85
    many = table.multiply_many(elems)
    naive = cref.batch_mul_g1(np.tile(g, len(elems)), h.scalars_bytes(elems))
    assert h.affine_g1(many) == h.affine_g1(naive)
    assert h.affine_g1(table.multiply(elems[0])) == h.affine_g1(naive[:144])
    assert h.affine_g1(table * elems[1]) == h.affine_g1(naive[144:288])
    table.free()
    out = msm.multiply_field_elems_with_same_group_elem(g, elems)
    assert h.affine_g1(out) == h.affine_g1(naive)


def test_variable_base_msm_api(dg, cref):
    bases, ks = h.g1_bases(64, 3)
    ss = h.rand_scalars(64, 4)
    v = msm.VariableBaseMSM()
    assert h.affine_g1(v.msm(bases, ss)) == h.known_dlog_msm_g1(ks, ss)
    with pytest.raises(msm.LengthMismatch) as ei:
        v.msm(bases, ss[:32 * 60])
    assert ei.value.min_len == 60
    # msm_unchecked / msm_bigint truncate
    assert h.affine_g1(v.msm_bigint(bases, ss[:32 * 60])) == h.known_dlog_msm_g1(ks[:32 * 60], ss[:32 * 60])
    assert h.affine_g1(msm.Pairs(bases, ss).msm()) == h.known_dlog_msm_g1(ks, ss)
    # 2-term MSM == g1*e1 + g2*e2 (utils/src/msm.rs:186-193)
    two = v.msm(bases[:192], ss[:64])
    naive = cref.batch_mul_g1(bases[:192], ss[:64])
    assert h.affine_g1(two) == h.affine_g1(dg.fold(naive))


def test_randomized_mult_checker(dg, cref):
    g, _ = h.g1_bases(3, 7)
    pts = [bytes(g[96 * i:96 * i + 96]) for i in range(3)]
    a = h.ints_of(h.rand_scalars(6, 8))

    def mul(p, s):
        return h.affine_g1(cref.batch_mul_g1(p, h.scalars_bytes([s])))

    def add(p, q):
        return o.g1_to_bytes(o.E1.add(o.g1_from_bytes(p), o.g1_from_bytes(q)))

    ck = msm.RandomizedMultChecker.new(0x1234567)
    ck.add_1(pts[0], a[0], mul(pts[0], a[0]))
    ck.add_2(pts[0], a[1], pts[1], a[2], add(mul(pts[0], a[1]), mul(pts[1], a[2])))
    ck.add_3(pts[0], a[3], pts[1], a[4], pts[2], a[5],
             add(add(mul(pts[0], a[3]), mul(pts[1], a[4])), mul(pts[2], a[5])))
    ck.add_many(pts, a[:3], add(add(mul(pts[0], a[0]), mul(pts[1], a[1])), mul(pts[2], a[2])))
    assert ck.verify()
    bad = msm.RandomizedMultChecker.new(0x1234567)
    bad.add_1(pts[0], a[0], mul(pts[0], a[0]))
    bad.add_1(pts[1], a[1], mul(pts[1], a[2]))          # wrong target
    assert not bad.verify()


@pytest.mark.parametrize('lazy', [True, False])
def test_randomized_pairing_checker(dg, cref, lazy):
    """test_pairing_randomize: true equations accepted, swapped outputs rejected, lazy == eager."""
    n = 4
    a, _ = h.g1_bases(n, 11)
    b, _ = h.g2_bases(n, 12)
    A = [bytes(a[96 * i:96 * i + 96]) for i in range(n)]
    B = [bytes(b[192 * i:192 * i + 192]) for i in range(n)]
    out1 = pc.multi_pairing(A[:2], B[:2])
    out2 = pc.pairing(A[2], B[2])
    ck = pc.RandomizedPairingChecker.new(0xDEADBEEFCAFE, lazy)
    ck.add_multiple_sources_and_target(A[:2], B[:2], out1)
    ck.add_sources_and_target(A[2], B[2], out2)
    # e(A3, B3) == e(s*A3, s^-1*B3)
    s = 0x77777
    sinv = pow(s, -1, o.R)
    c = h.affine_g1(cref.batch_mul_g1(A[3], h.scalars_bytes([s])))
    d = h.affine_g2(cref.batch_mul_g2(B[3], h.scalars_bytes([sinv])))
    ck.add_sources(A[3], B[3], c, d)
    ck.add_multiple_sources([A[3], A[0]], [B[3], B[0]], [c, A[0]], [d, B[0]])
    assert ck.verify()
    bad = pc.RandomizedPairingChecker.new(0xDEADBEEFCAFE, lazy)
    bad.add_multiple_sources_and_target(A[:2], B[:2], out2)      # swapped outputs
    bad.add_sources_and_target(A[2], B[2], out1)
    assert not bad.verify()


def test_bbs_plus_shaped_sign_verify(dg, cref):
    """bbs_plus/src/signature.rs:138-211, 272-296 with synthetic parameters: b = g1 + h0*s +
    sum h_i m_i via ONE MSM; A = b * 1/(e+x); verify e(A, pk + g2*e) == e(b, g2), i.e. the
    2-pair product check e(A, pk + e*g2) * e(-b, g2) == 1; a tampered message must fail."""
    n = 200
    hs, hk = h.g1_bases(n + 2, 31)                   # g1, h0, h_1..h_n
    msgs = h.ints_of(h.rand_scalars(n, 32))
    s, e, x = 0x1111, 0x2222, 0x3333
    scal = h.scalars_bytes([1, s] + msgs)
    b_jac = msm.VariableBaseMSM().msm_unchecked(hs, scal)
    b_aff = msm.into_affine(b_jac)
    assert b_aff == h.known_dlog_msm_g1(hk, scal)
    A = h.affine_g1(dg.batch_mul(b_aff, h.scalars_bytes([pow(e + x, -1, o.R)])))
    g2 = o.g2_to_bytes(o.G2_GEN)
    pk_plus = bytes(cref.g2_generator_muls(h.scalars_bytes([(x + e) % o.R])))
    assert dg.multi_pairing_is_one(A + h.neg_g1(b_aff), pk_plus + g2)
    msgs[5] += 1
    b_bad = msm.into_affine(msm.VariableBaseMSM().msm_unchecked(hs, h.scalars_bytes([1, s] + msgs)))
    assert not dg.multi_pairing_is_one(A + h.neg_g1(b_bad), pk_plus + g2)
