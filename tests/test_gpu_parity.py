"""GPU parity tests (run with -m gpu on the B200 box): every result that crosses the C ABI is
compared bit-for-bit (canonical affine / Fp12 bytes) with the CPU oracle on the same seeded
inputs, with the committed big-int golden vectors, and - at full benchmark sizes - through the
O(n) known-discrete-log identity  sum s_i (k_i G) = (sum s_i k_i) G.

The relational checks mirror the reference's own tests: utils/src/msm.rs:116-308,
utils/src/randomized_pairing_check.rs:234-421, utils/src/randomized_mult_checker.rs:136-384."""
import json
import os

import numpy as np
import pytest

from oracle import bls12_381 as o
from tests import helpers as h

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'vectors.json')))


# ------------------------------------------------------------------ field --------------------
@pytest.mark.parametrize('op', [0, 1, 2, 3, 4])
def test_fp_ops_vs_bigint(dg, op):
    n = 512
    rng = np.random.default_rng(op)
    xs = [int.from_bytes(rng.bytes(48), 'little') % o.P for _ in range(n)]
    ys = [int.from_bytes(rng.bytes(48), 'little') % o.P for _ in range(n)]
    xs[:4] = [0, 1, o.P - 1, o.P - 2]
    ys[:4] = [0, o.P - 1, o.P - 1, 1]
    a = b''.join(o.fp_to_mont_bytes(x) for x in xs)
    b = b''.join(o.fp_to_mont_bytes(y) for y in ys)
    out = bytes(dg.dbg_fp_op(op, a, b))
    f = {0: lambda x, y: x * y, 1: lambda x, y: x + y, 2: lambda x, y: x - y, 3: lambda x, y: x * x,
         4: lambda x, y: -x}[op]
    exp = b''.join(o.fp_to_mont_bytes(f(x, y) % o.P) for x, y in zip(xs, ys))
    assert out == exp


def test_fp_sqr_dedicated(dg):
    """The dedicated squaring (66 cross products + 12 squares + 12 reduction rows, fp.cuh) against Python integers, on raw
    Montgomery representations chosen for their limb patterns (all-ones limbs, single bits, p - 1) and on random values;
    the same inputs through the general multiplier must agree bit for bit."""
    rng = np.random.default_rng(33)
    rinv = pow(1 << 384, -1, o.P)
    raws = [0, 1, 2, o.P - 1, o.P - 2, (1 << 380) - 1, (1 << 380), (1 << 380) + (1 << 379) - 1, (1 << 352) - 1, ((1 << 28) - 1) << 352,
            0xffffffff, 0xffffffff00000000, (1 << 64) - 1, ((1 << 380) - 1) ^ ((1 << 190) - 1), int('55' * 47, 16), int('aa' * 47, 16) >> 3]
    raws += [1 << k for k in range(0, 381, 5)] + [o.P - (1 << k) for k in range(0, 380, 9)]
    raws += [((1 << 32) - 1) << (32 * k) for k in range(11)] + [(1 << 380) - 1 - (((1 << 32) - 1) << (32 * k)) for k in range(11)]
    raws += [int.from_bytes(rng.bytes(48), 'little') % o.P for _ in range(4096)]
    assert all(0 <= v < o.P for v in raws)
    a = b''.join(v.to_bytes(48, 'little') for v in raws)
    out = bytes(dg.dbg_fp_op(3, a, a))
    exp = b''.join(o.fp_to_mont_bytes(pow(v * rinv % o.P, 2, o.P)) for v in raws)
    assert out == exp
    assert bytes(dg.dbg_fp_op(0, a, a)) == exp


@pytest.mark.parametrize('op', [5, 6, 7])       # 5: Fermat, 6: bit-serial binary Euclid, 7: Pornin's binary GCD (31-bit inner rounds)
def test_fp_inverse(dg, op):
    rng = np.random.default_rng(7)
    xs = [1, 2, o.P - 1, 0x1234567890ABCDEF, 0, 3, (o.P + 1) // 2, o.P - 2, (o.P - 1) // 2, 1 << 380, (1 << 381) % o.P, (1 << 64) - 1,
          1 << 64, (1 << 96) + 1, 0xffffffff, 1 << 31, (1 << 31) - 1, 1 << 32, 1 << 33, (1 << 381) - 1 - o.P]
    xs += [1 << k for k in range(0, 381, 7)] + [(o.P - (1 << k)) % o.P for k in range(0, 381, 11)]
    xs += [int.from_bytes(rng.bytes(48), 'little') % o.P for _ in range(2000)]
    xs += [int.from_bytes(rng.bytes(k), 'little') for k in (1, 4, 8, 9, 16, 24, 40)]
    a = b''.join(o.fp_to_mont_bytes(x) for x in xs)
    out = bytes(dg.dbg_fp_op(op, a, a))
    exp = b''.join(o.fp_to_mont_bytes(pow(x, o.P - 2, o.P)) for x in xs)
    assert out == exp


# ------------------------------------------------------------------ MSM ----------------------
def test_msm_g1_golden(dg):
    g = GOLD['msm_g1']
    res = dg.msm(bytes.fromhex(g['bases']), bytes.fromhex(g['scalars']))
    assert h.affine_g1(res).hex() == g['result_affine']


def test_msm_g2_golden(dg):
    g = GOLD['msm_g2']
    res = dg.msm(bytes.fromhex(g['bases']), bytes.fromhex(g['scalars']), g2=True)
    assert h.affine_g2(res).hex() == g['result_affine']


@pytest.mark.parametrize('n', [0, 1, 2, 31, 32, 33, 100, 1000, 4097, 10001])
def test_msm_g1_vs_oracle(dg, cref, n):
    bases, ks = h.g1_bases(n, 1000 + n)
    ss = h.rand_scalars(n, 2000 + n)
    res = dg.msm(bases, ss, n=n)
    assert h.affine_g1(res) == h.affine_g1(cref.msm_g1(bases, ss, n))
    if n:
        assert h.affine_g1(res) == h.known_dlog_msm_g1(ks, ss)


@pytest.mark.parametrize('n', [0, 1, 33, 1000])
def test_msm_g2_vs_oracle(dg, cref, n):
    bases, ks = h.g2_bases(n, 3000 + n)
    ss = h.rand_scalars(n, 4000 + n)
    res = dg.msm(bases, ss, g2=True, n=n)
    assert h.affine_g2(res) == h.affine_g2(cref.msm_g2(bases, ss, n))
    if n:
        assert h.affine_g2(res) == h.known_dlog_msm_g2(ks, ss)


@pytest.mark.parametrize('c', [4, 7, 8, 11, 13, 16])
def test_msm_g1_every_window_size(dg, cref, c):
    n = 3000
    bases, ks = h.g1_bases(n, 77)
    ss = h.rand_scalars(n, 78)
    dg.msm_set_window(c)
    try:
        res = dg.msm(bases, ss)
    finally:
        dg.msm_set_window(0)
    assert h.affine_g1(res) == h.known_dlog_msm_g1(ks, ss)


def test_msm_g1_adversarial_inputs(dg, cref):
    """Edge cases arkworks handles implicitly (SURVEY.md section 7 'hard parts'): identity bases,
    zero / one / r-1 scalars, all-equal bases, P and -P meeting in one bucket, all-equal scalars
    (one hot bucket per window), tiny scalars as in real witnesses."""
    n = 2048
    bases, ks = h.g1_bases(n, 5)
    bases = bases.copy()
    ss = h.rand_scalars(n, 6).copy()
    sv = ss.reshape(n, 32)
    bv = bases.reshape(n, 96)
    bv[10] = 0; bv[11] = 0                                  # identity bases
    sv[20] = 0; sv[21] = 0; sv[21, 0] = 1                   # scalars 0 and 1
    sv[22] = np.frombuffer((o.R - 1).to_bytes(32, 'little'), np.uint8)
    bv[100:200] = bv[100]                                   # 100 equal bases
    sv[100:150] = sv[100]                                   # ... half of them with equal scalars
    bv[301] = np.frombuffer(h.neg_g1(bv[300]), np.uint8)    # P, -P with the same scalar
    sv[301] = sv[300]
    sv[400:900] = 0; sv[400:900, 0] = 1                     # 500 scalars equal to one
    sv[900:1200, 2:] = 0                                    # 16-bit scalars
    res = dg.msm(bases, ss)
    assert h.affine_g1(res) == h.affine_g1(cref.msm_g1(bases, ss))
    # everything cancels -> identity
    two = np.concatenate([bv[300], bv[301]])
    res = dg.msm(two, np.concatenate([sv[300], sv[300]]))
    assert not any(h.affine_g1(res))


def test_msm_g1_all_scalars_equal_is_balanced(dg, cref):
    n = 1 << 14
    bases, ks = h.g1_bases(n, 9)
    ss = np.tile(h.rand_scalars(1, 10), n)
    res = dg.msm(bases, ss)
    assert h.affine_g1(res) == h.known_dlog_msm_g1(ks, ss)


def test_msm_rejects_non_canonical_scalar(dg):
    bases, _ = h.g1_bases(4, 1)
    ss = h.rand_scalars(4, 2).copy()
    ss[31] = 0xFF                                            # >= 2^255
    with pytest.raises(dg.DockGpuError):
        dg.msm(bases, ss)
    good = h.rand_scalars(4, 2)
    dg.msm(bases, good)                                      # library still usable afterwards


def test_msm_resident_bases_handle_and_truncation(dg, cref):
    n = 5000
    bases, ks = h.g1_bases(n, 21)
    ss = h.rand_scalars(n, 22)
    hb = dg.Bases(bases)
    try:
        assert h.affine_g1(dg.msm(hb, ss)) == h.known_dlog_msm_g1(ks, ss)
        # fewer scalars than bases: truncate like ark msm_bigint
        assert h.affine_g1(dg.msm(hb, ss[:32 * 100])) == h.known_dlog_msm_g1(ks[:32 * 100], ss[:32 * 100])
    finally:
        hb.free()


@pytest.mark.parametrize('logn', [16, 18, 20])
def test_msm_g1_benchmark_sizes_known_dlog(dg, logn):
    """BASELINE.json sizes through the size-independent known-dlog identity."""
    n = 1 << logn
    bases, ks = h.g1_bases(n, 0xD0C4C0DE ^ n)
    ss = h.rand_scalars(n, 0xC0FFEE ^ n)
    assert h.affine_g1(dg.msm(bases, ss)) == h.known_dlog_msm_g1(ks, ss)


def test_msm_linearity(dg):
    """MSM(P, s) + MSM(P, t) == MSM(P, s + t)."""
    n = 3000
    bases, _ = h.g1_bases(n, 31)
    s, t = h.ints_of(h.rand_scalars(n, 32)), h.ints_of(h.rand_scalars(n, 33))
    a = dg.msm(bases, h.scalars_bytes(s)); b = dg.msm(bases, h.scalars_bytes(t))
    c = dg.msm(bases, h.scalars_bytes([(x + y) % o.R for x, y in zip(s, t)]))
    assert h.affine_g1(dg.fold(np.concatenate([a, b]))) == h.affine_g1(c)


# ------------------------------------------------------------------ batch ops ----------------
@pytest.mark.parametrize('hint_n,m', [(10, 10), (30, 30), (10000, 300), (1 << 19, 50)])
def test_fixed_base_vs_oracle(dg, cref, hint_n, m):
    """utils/src/msm.rs:213-229, 284-307: WindowTable multiply == per-scalar mul_bigint."""
    base, _ = h.g1_bases(1, 50 + m)
    ss = h.rand_scalars(m, 60 + m).copy()
    ss[:32] = 0
    t = dg.FixedBaseTable(base, hint_n)
    try:
        exp, window, nwin = cref.fixed_base_mul_many_g1(base, hint_n, ss)
        assert (t.window, t.num_windows) == (window, nwin)
        assert h.affine_g1(t.mul_many(ss)) == h.affine_g1(exp)
        assert h.affine_g1(t.mul_many(ss)) == h.affine_g1(cref.batch_mul_g1(np.tile(base, m), ss))
        if hint_n <= 10000:
            tbl = t.download()
            assert bytes(tbl) == bytes(cref.fixed_base_table_g1(base, window))
    finally:
        t.free()


def test_fixed_base_golden_and_g2(dg, cref):
    g = GOLD['fixed_base_g1']
    t = dg.FixedBaseTable(bytes.fromhex(g['point']), 7)
    assert h.affine_g1(t.mul_many(bytes.fromhex(g['scalars']))).hex() == g['results_affine']
    t.free()
    base, _ = h.g2_bases(1, 5)
    ss = h.rand_scalars(40, 6)
    t2 = dg.FixedBaseTable(base, 40, g2=True)
    exp, _, _ = cref.fixed_base_mul_many_g2(base, 40, ss)
    assert h.affine_g2(t2.mul_many(ss)) == h.affine_g2(exp)
    t2.free()


def test_batch_mul_vs_oracle(dg, cref):
    m = 257
    pts, _ = h.g1_bases(m, 70)
    pts = pts.copy()
    ss = h.rand_scalars(m, 71).copy()
    pts[:96] = 0                                  # identity point
    ss[32:64] = 0                                 # zero scalar
    assert h.affine_g1(dg.batch_mul(pts, ss)) == h.affine_g1(cref.batch_mul_g1(pts, ss))
    p2, _ = h.g2_bases(33, 72)
    s2 = h.rand_scalars(33, 73)
    assert h.affine_g2(dg.batch_mul(p2, s2, g2=True)) == h.affine_g2(cref.batch_mul_g2(p2, s2))


def test_normalize_batch_vs_oracle(dg, cref):
    m = 100
    pts, _ = h.g1_bases(m, 80)
    jac = cref.batch_mul_g1(pts, h.rand_scalars(m, 81)).copy()
    jac[144 * 5 + 96:144 * 6] = 0                 # z = 0 -> identity, must be skipped by the batch inversion
    assert bytes(dg.normalize_batch(jac)) == bytes(cref.normalize_batch_g1(jac))
    p2, _ = h.g2_bases(20, 82)
    j2 = cref.batch_mul_g2(p2, h.rand_scalars(20, 83))
    assert bytes(dg.normalize_batch(j2, g2=True)) == bytes(cref.normalize_batch_g2(j2))


def test_fused_witness_update(dg, cref):
    """vb_accumulator/src/witness.rs:269-284: new_C_i = a_i * C_i + b_i * V, normalised."""
    m = 200
    pts, _ = h.g1_bases(m, 90)
    v, _ = h.g1_bases(1, 91)
    sa, sb = h.rand_scalars(m, 92), h.rand_scalars(m, 93)
    sa, sb = np.array(sa), np.array(sb)
    R_ = o.R
    # corner cases: zero scalars, a = b = r - 1, an integer >= r (generic chain), an identity witness, C_i = +-V
    sa[:32] = 0
    sb[32:64] = 0
    sa[64:96] = np.frombuffer((R_ - 1).to_bytes(32, 'little'), dtype=np.uint8)
    sb[96:128] = np.frombuffer((R_ + 5).to_bytes(32, 'little'), dtype=np.uint8)
    pts = np.array(pts)
    pts[96 * 4:96 * 5] = 0
    pts[96 * 5:96 * 6] = np.frombuffer(bytes(v), dtype=np.uint8)
    pts[96 * 6:96 * 7] = np.frombuffer(h.neg_g1(bytes(v)), dtype=np.uint8)
    t = dg.FixedBaseTable(v, m)
    outs = []
    try:
        # 1, 6: two threads per element / 3: one joint doubling chain per thread / 2: window table /
        # 4: one quad per product (the default at this size) / 5: no quads
        for force in (1, 3, 2, 6, 4, 5):
            dg.dbg_set_tunable(5, force)
            outs.append(bytes(dg.batch_mul_add_fixed_g1(pts, sa, t, sb)))
            if force in (4, 5, 6):
                outs.append(bytes(dg.batch_mul_add_same_g1(pts, sa, v, sb)))
    finally:
        dg.dbg_set_tunable(5, 0)
    t.free()
    outs.append(bytes(dg.batch_mul_add_same_g1(pts, sa, v, sb)))
    assert all(x == outs[0] for x in outs)
    # batch_mul itself: quad kernel == thread-per-element kernel (Jacobian outputs differ in representation, not in value)
    try:
        dg.dbg_set_tunable(5, 4)
        bq = bytes(dg.normalize_batch(dg.batch_mul(pts, sb)))
        dg.dbg_set_tunable(5, 5)
        bt = bytes(dg.normalize_batch(dg.batch_mul(pts, sb)))
    finally:
        dg.dbg_set_tunable(5, 0)
    assert bq == bt == bytes(cref.normalize_batch_g1(cref.batch_mul_g1(pts, sb)))
    out = outs[0]
    # V = identity: only the a_i * C_i terms remain
    only_a = bytes(dg.batch_mul_add_same_g1(pts, sa, bytes(96), sb))
    assert only_a == bytes(dg.normalize_batch(dg.batch_mul(pts, sa)))
    left = cref.batch_mul_g1(pts, sa)
    right, _, _ = cref.fixed_base_mul_many_g1(v, m, sb)
    exp = b''
    for i in range(m):
        a = o.g1_from_bytes(h.affine_g1(left[144 * i:144 * i + 144]))
        b = o.g1_from_bytes(h.affine_g1(right[144 * i:144 * i + 144]))
        exp += o.g1_to_bytes(o.E1.add(a, b))
    assert bytes(out) == exp


def test_fold(dg, cref):
    pts, ks = h.g1_bases(8, 95)
    jac = cref.batch_mul_g1(pts, h.scalars_bytes([1] * 8))
    tot = sum(h.ints_of(ks)) % o.R
    assert h.affine_g1(dg.fold(jac)) == bytes(cref.g1_generator_muls(h.scalars_bytes([tot])))


# ------------------------------------------------------------------ pairing -------------------
def test_pairing_golden(dg):
    g = GOLD['pairing']
    gen1, gen2 = bytes.fromhex(GOLD['g1_generator']), bytes.fromhex(GOLD['g2_generator'])
    assert bytes(dg.multi_pairing(gen1, gen2)).hex() == g['e_g1_g2']
    assert bytes(dg.multi_pairing(bytes.fromhex(g['p']), bytes.fromhex(g['q']))).hex() == g['e_p_q']


def test_miller_loop_and_final_exp_vs_oracle(dg, cref):
    k = 5
    ps, _ = h.g1_bases(k, 110)
    qs, _ = h.g2_bases(k, 111)
    ps = ps.copy(); ps[96:192] = 0                # one identity pair is dropped
    ml = dg.multi_miller_loop(ps, qs)
    assert bytes(ml) == bytes(cref.multi_miller_loop(ps, qs))
    fe = dg.final_exponentiation(ml)
    assert bytes(fe) == bytes(cref.final_exp(ml))
    assert bytes(dg.multi_pairing(ps, qs)) == bytes(fe)
    assert dg.final_exponentiation(bytes(576)) is None
    assert bytes(dg.multi_pairing(b'', b'')) == bytes(cref.fp12_one())


def test_pairing_bilinearity_and_product_check(dg, cref):
    """e(aP, bQ) == e(P, Q)^(ab); e(P,Q) e(-P,Q) == 1 (the verifier's product check)."""
    a, b = 0x1234567, 0xABCDEF123
    gen1, gen2 = bytes.fromhex(GOLD['g1_generator']), bytes.fromhex(GOLD['g2_generator'])
    pa = bytes(cref.g1_generator_muls(h.scalars_bytes([a])))
    qb = bytes(cref.g2_generator_muls(h.scalars_bytes([b])))
    e_ab = dg.multi_pairing(pa, qb)
    e0 = dg.multi_pairing(gen1, gen2)
    assert bytes(dg.gt_pow(e0, h.scalars_bytes([a * b % o.R]))) == bytes(e_ab)
    assert bytes(dg.gt_pow(e0, h.scalars_bytes([a * b % o.R]))) == bytes(cref.fp12_pow(e0, h.scalars_bytes([a * b % o.R])))
    assert dg.multi_pairing_is_one(pa + h.neg_g1(pa), qb + qb)
    assert not dg.multi_pairing_is_one(pa + pa, qb + qb)


def test_many_pairs_tree_product(dg, cref):
    k = 70                                        # exercises two levels of the CTA product tree
    ps, _ = h.g1_bases(k, 120)
    qs, _ = h.g2_bases(k, 121)
    assert bytes(dg.multi_miller_loop(ps, qs)) == bytes(cref.multi_miller_loop(ps, qs))


# ------------------------------------------------------------------ precomputed bases ----------
@pytest.mark.parametrize('c', [8, 13, 16, 20])
def test_msm_precomputed_bases(dg, cref, c):
    """dg_bases_precompute folds all digit positions into one bucket set; same group element."""
    n = 3000
    bases, ks = h.g1_bases(n, 300 + c)
    bases = bases.copy(); bases[96 * 7:96 * 8] = 0           # an identity base must stay harmless
    ss = h.rand_scalars(n, 400 + c).copy()
    ss[:32] = 0
    ss[32:64] = np.frombuffer((o.R - 1).to_bytes(32, 'little'), np.uint8)
    hb = dg.Bases(bases).precompute(c)
    try:
        exp = h.affine_g1(cref.msm_g1(bases, ss))
        assert h.affine_g1(dg.msm(hb, ss)) == exp
        # prefix use of a precomputed table
        assert h.affine_g1(dg.msm(hb, ss[:32 * 1000])) == h.affine_g1(cref.msm_g1(bases, ss[:32 * 1000], 1000))
        with pytest.raises(dg.DockGpuError):
            hb.precompute(c)                                   # already precomputed
    finally:
        hb.free()


def test_msm_precomputed_g2_and_default_window(dg, cref):
    n = 600
    bases, ks = h.g2_bases(n, 501)
    ss = h.rand_scalars(n, 502)
    hb = dg.Bases(bases, g2=True).precompute()
    try:
        assert h.affine_g2(dg.msm(hb, ss, g2=True)) == h.known_dlog_msm_g2(ks, ss)
    finally:
        hb.free()


def test_msm_unchecked_montgomery_scalars(dg, cref):
    """msm_unchecked: Fr scalars in ark-ff's Montgomery form are converted on the device (into_bigint)."""
    n = 777
    bases, ks = h.g1_bases(n, 601)
    ss = h.rand_scalars(n, 602)
    mont = b''.join(o.fr_to_mont_bytes(v) for v in h.ints_of(ss))
    assert bytes(dg.fr_into_bigint(mont)) == bytes(ss)
    assert h.affine_g1(dg.msm_unchecked(bases, mont)) == h.known_dlog_msm_g1(ks, ss)
    b2, k2 = h.g2_bases(50, 603)
    assert h.affine_g2(dg.msm_unchecked(b2, mont[:32 * 50], g2=True)) == h.known_dlog_msm_g2(k2, ss[:32 * 50])


def test_multi_pairing_batch(dg, cref):
    ps, _ = h.g1_bases(30, 130)
    qs, _ = h.g2_bases(30, 131)
    ps = ps.copy(); ps[96 * 3:96 * 4] = 0
    counts = [1, 9, 0, 12, 8]
    outs = dg.multi_pairing_batch(ps, qs, counts)
    off = 0
    for c, got in zip(counts, outs):
        exp = cref.multi_pairing(ps[96 * off:96 * (off + c)], qs[192 * off:192 * (off + c)]) if c else cref.fp12_one()
        assert bytes(got) == bytes(exp)
        off += c
