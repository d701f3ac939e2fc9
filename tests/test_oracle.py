"""CPU: pin the oracle.  The reference holds no golden vectors for this path ("parity
unpinned"), so the oracle is pinned on public constants, on an independent
second derivation of the pairing, on the committed big-int golden vectors, and on the
relational checks the reference's own tests use (utils/src/msm.rs:116-308)."""
import json
import os

import numpy as np
import pytest

from oracle import bls12_381 as o
from tests import helpers as h

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'vectors.json')))


def test_public_constants():
    assert o.E1.on_curve(o.G1_GEN) and o.E2.on_curve(o.G2_GEN)
    assert o.E1.mul(o.G1_GEN, o.R) is None
    assert o.E2.mul(o.G2_GEN, o.R) is None
    # the Zcash / IETF compressed generator every BLS12-381 library agrees on
    assert o.g1_compressed(o.G1_GEN).hex() == GOLD['g1_generator_compressed']
    assert GOLD['g1_generator_compressed'].startswith('97f1d3a73197d7942695638c4fa9ac0f')


def test_ark_window_rules():
    # c = 13 @2^16, 14 @2^18, 15 @2^20, 17 @2^22, 18 @2^24
    assert [o.msm_window_size(1 << k) for k in (16, 18, 20, 22, 24)] == [13, 14, 15, 17, 18]
    assert o.msm_window_size(31) == 3
    assert o.fixed_base_window_size(10) == 3 and o.fixed_base_window_size(10000) == 9


def test_signed_digits_reconstruct():
    rng = o.SplitMix64(5)
    for c in (3, 7, 13, 15, 16):
        for _ in range(20):
            s = rng.scalar()
            d = o.make_digits(s, c)
            assert sum(x << (c * i) for i, x in enumerate(d)) == s
            assert all(-(1 << (c - 1)) <= x < (1 << (c - 1)) for x in d[:-1])


def test_pairing_two_derivations_agree():
    """ark-style tower algorithm == textbook polynomial-ring pairing cubed."""
    e_ark = o.pairing(o.G1_GEN, o.G2_GEN)
    assert o.fp12_to_bytes(e_ark).hex() == GOLD['pairing']['e_g1_g2']
    assert e_ark != o.FP12_ONE
    assert o.fp12_pow(e_ark, o.R) == o.FP12_ONE


def test_pairing_bilinear():
    a, b = 0x1234567, 0xABCDEF123
    e0 = o.pairing(o.G1_GEN, o.G2_GEN)
    e1 = o.pairing(o.E1.mul(o.G1_GEN, a), o.E2.mul(o.G2_GEN, b))
    assert o.fp12_pow(e0, a * b % o.R) == e1


def test_c_oracle_field_and_generators(cref):
    ks = h.scalars_bytes([1, 2, o.R - 1])
    g = cref.g1_generator_muls(ks)
    assert bytes(g[:96]).hex() == GOLD['g1_generator']
    assert bytes(g[96:192]) == o.g1_to_bytes(o.E1.mul(o.G1_GEN, 2))
    assert bytes(g[192:288]) == o.g1_to_bytes(o.E1.neg(o.G1_GEN))
    g2 = cref.g2_generator_muls(ks)
    assert bytes(g2[:192]).hex() == GOLD['g2_generator']


def test_c_oracle_msm_golden(cref):
    g = GOLD['msm_g1']
    res = cref.msm_g1(bytes.fromhex(g['bases']), bytes.fromhex(g['scalars']))
    assert h.affine_g1(res).hex() == g['result_affine']
    g = GOLD['msm_g2']
    res = cref.msm_g2(bytes.fromhex(g['bases']), bytes.fromhex(g['scalars']))
    assert h.affine_g2(res).hex() == g['result_affine']


@pytest.mark.parametrize('n', [0, 1, 2, 31, 32, 33, 200])
def test_c_oracle_msm_vs_known_dlog(cref, n):
    bases, ks = h.g1_bases(n, 100 + n)
    ss = h.rand_scalars(n, 200 + n)
    res = cref.msm_g1(bases, ss, n)
    if n == 0:
        assert not h.affine_g1(res).strip(b'\0')
    else:
        assert h.affine_g1(res) == h.known_dlog_msm_g1(ks, ss)


def test_c_oracle_msm_matches_python_pippenger(cref):
    n = 40
    bases, ks = h.g1_bases(n, 1)
    ss = h.rand_scalars(n, 2)
    pts = [o.g1_from_bytes(bytes(bases[96 * i:96 * i + 96])) for i in range(n)]
    exp = o.msm_pippenger(o.E1, pts, h.ints_of(ss))
    assert h.affine_g1(cref.msm_g1(bases, ss)) == o.g1_to_bytes(exp)
    assert exp == o.E1.msm_naive(pts, h.ints_of(ss))


def test_c_oracle_msm_truncates_to_shorter(cref):
    bases, ks = h.g1_bases(10, 3)
    ss = h.rand_scalars(7, 4)
    assert h.affine_g1(cref.msm_g1(bases, ss)) == h.known_dlog_msm_g1(ks[:7 * 32], ss)


def test_c_oracle_fixed_base_golden(cref):
    g = GOLD['fixed_base_g1']
    out, window, nwin = cref.fixed_base_mul_many_g1(bytes.fromhex(g['point']), 7, bytes.fromhex(g['scalars']))
    assert (window, nwin) == (3, 85)
    assert h.affine_g1(out).hex() == g['results_affine']
    # larger hint -> different window, same group elements (utils/src/msm.rs:296-303)
    out2, window2, _ = cref.fixed_base_mul_many_g1(bytes.fromhex(g['point']), 10000, bytes.fromhex(g['scalars']))
    assert window2 == 9 and h.affine_g1(out2).hex() == g['results_affine']


def test_c_oracle_batch_mul_and_normalize(cref):
    n = 9
    bases, ks = h.g1_bases(n, 7)
    ss = h.rand_scalars(n, 8)
    out = cref.batch_mul_g1(bases, ss)
    aff = h.affine_g1(out)
    for i, (k, s) in enumerate(zip(h.ints_of(ks), h.ints_of(ss))):
        assert aff[96 * i:96 * i + 96] == o.g1_to_bytes(o.E1.mul(o.G1_GEN, k * s % o.R))


def test_c_oracle_pairing_golden(cref):
    g = GOLD['pairing']
    gen1, gen2 = bytes.fromhex(GOLD['g1_generator']), bytes.fromhex(GOLD['g2_generator'])
    assert bytes(cref.multi_pairing(gen1, gen2)).hex() == g['e_g1_g2']
    p, q = bytes.fromhex(g['p']), bytes.fromhex(g['q'])
    assert bytes(cref.multi_pairing(p, q)).hex() == g['e_p_q']
    # e(P,Q) * e(-P,Q) == 1, and identity pairs are skipped
    one = bytes(cref.fp12_one())
    assert bytes(cref.multi_pairing(p + bytes.fromhex(g['neg_p']), q + q)) == one
    assert bytes(cref.multi_pairing(p + bytes(96), q + q)).hex() == g['e_p_q']
    assert cref.final_exp(bytes(576)) is None


def test_c_oracle_miller_matches_python(cref):
    g = GOLD['pairing']
    p, q = bytes.fromhex(g['p']), bytes.fromhex(g['q'])
    ml = o.multi_miller_loop([o.g1_from_bytes(p)], [o.g2_from_bytes(q)])
    assert bytes(cref.multi_miller_loop(p, q)) == o.fp12_to_bytes(ml)


# ------------------------------------------------------------------ Fr NTT (row f1) -----------
def test_fr_root_of_unity_is_arkworks_constant():
    assert o.FR_TWO_ADIC_ROOT == 10238227357739495823651030575849232062558860180284477541189508159991286009131
    g = o.fr_domain_generator(19)
    assert pow(g, 1 << 19, o.R) == 1 and pow(g, 1 << 18, o.R) != 1


@pytest.mark.parametrize('logn', [0, 1, 3, 6])
def test_c_oracle_ntt_vs_definition(cref, logn):
    n = 1 << logn
    rng = o.SplitMix64(logn + 1)
    co = [rng.scalar() for _ in range(n)]
    data = b''.join(o.fr_to_mont_bytes(c) for c in co)
    for coset in (False, True):
        ev = cref.fr_ntt(data, logn, False, coset)
        assert bytes(ev) == b''.join(o.fr_to_mont_bytes(v) for v in o.fr_fft_definition(co, logn, coset))
        assert bytes(cref.fr_ntt(ev, logn, True, coset)) == data


def test_c_oracle_ntt_convolution(cref):
    """ifft(fft(a) * fft(b)) is the cyclic convolution: pins the transform through polynomial products."""
    logn, n = 5, 32
    rng = o.SplitMix64(9)
    a = [rng.scalar() for _ in range(n // 2)] + [0] * (n // 2)
    b = [rng.scalar() for _ in range(n // 2)] + [0] * (n // 2)
    enc = lambda v: b''.join(o.fr_to_mont_bytes(x) for x in v)
    fa = [o.fr_from_mont_bytes(bytes(cref.fr_ntt(enc(a), logn)[32 * i:32 * i + 32])) for i in range(n)]
    fb = [o.fr_from_mont_bytes(bytes(cref.fr_ntt(enc(b), logn)[32 * i:32 * i + 32])) for i in range(n)]
    prod = cref.fr_ntt(enc([x * y % o.R for x, y in zip(fa, fb)]), logn, True)
    exp = [0] * n
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            exp[(i + j) % n] = (exp[(i + j) % n] + x * y) % o.R
    assert bytes(prod) == enc(exp)
