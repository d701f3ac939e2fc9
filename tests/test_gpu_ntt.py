"""GPU parity for the Fr NTT: ark_poly Radix2EvaluationDomain fft / ifft and
the coset variants (offset 7), and the witness-map tail of legogroth16/src/r1cs_to_qap.rs:187-207.
Bit-exact against the C oracle (itself pinned to the big-int definition in tests/test_oracle.py);
at full size through the size-independent round trip and the convolution theorem."""
import numpy as np
import pytest

from oracle import bls12_381 as o
from tests import helpers as h

pytestmark = pytest.mark.gpu


def mont_fr(n, seed):
    """n random Fr elements as Montgomery records (any canonical value < r is a valid record)."""
    return h.rand_scalars(n, seed)


@pytest.mark.parametrize('op', [0, 1, 2, 3])
def test_fr_ops_vs_bigint(dg, op):
    n = 256
    xs = h.ints_of(h.rand_scalars(n, 31)); ys = h.ints_of(h.rand_scalars(n, 32))
    xs[:4] = [0, 1, o.R - 1, o.R - 2]; ys[:4] = [0, o.R - 1, o.R - 1, 1]
    enc = lambda v: b''.join(o.fr_to_mont_bytes(x) for x in v)
    f = {0: lambda x, y: x * y, 1: lambda x, y: x + y, 2: lambda x, y: x - y, 3: lambda x, y: pow(x, o.R - 2, o.R)}[op]
    assert bytes(dg.dbg_fr_op(op, enc(xs), enc(ys))) == enc([f(x, y) % o.R for x, y in zip(xs, ys)])


@pytest.mark.parametrize('logn', [0, 1, 2, 5, 8, 9, 11, 14, 15, 18])
@pytest.mark.parametrize('inverse,coset', [(False, False), (True, False), (False, True), (True, True)])
def test_ntt_vs_oracle(dg, cref, logn, inverse, coset):
    data = mont_fr(1 << logn, 700 + logn)
    got = dg.fr_ntt(data, logn, inverse, coset)
    exp = cref.fr_ntt(data, logn, inverse, coset)
    assert np.array_equal(got, exp)


def test_ntt_matches_bigint_definition(dg):
    logn = 4
    coeffs = h.ints_of(h.rand_scalars(1 << logn, 5))
    data = b''.join(o.fr_to_mont_bytes(c) for c in coeffs)
    for coset in (False, True):
        got = bytes(dg.fr_ntt(data, logn, False, coset))
        assert got == b''.join(o.fr_to_mont_bytes(v) for v in o.fr_fft_definition(coeffs, logn, coset))


@pytest.mark.parametrize('logn', [19, 20])
def test_ntt_round_trip_full_size(dg, logn):
    data = mont_fr(1 << logn, 800 + logn)
    for coset in (False, True):
        ev = dg.fr_ntt(data, logn, False, coset)
        assert not np.array_equal(ev, data)
        assert np.array_equal(dg.fr_ntt(ev, logn, True, coset), data)


def test_ntt_linearity_and_delta(dg):
    logn = 12
    n = 1 << logn
    one = np.frombuffer(o.fr_to_mont_bytes(1), np.uint8)
    delta = np.zeros(32 * n, np.uint8); delta[:32] = one
    ev = dg.fr_ntt(delta, logn)                       # FFT of the constant polynomial 1: all ones
    assert np.array_equal(ev, np.tile(one, n))


def test_qap_h_from_abc(dg, cref):
    """legogroth16/src/r1cs_to_qap.rs:187-207 on synthetic a, b, c with c = a*b on the domain, so that
    (ab - c) is divisible by Z and h is a genuine quotient polynomial (degree < n - 1)."""
    logn = 10
    n = 1 << logn
    a = h.ints_of(h.rand_scalars(n, 1)); b = h.ints_of(h.rand_scalars(n, 2))
    c = [x * y % o.R for x, y in zip(a, b)]
    enc = lambda v: b''.join(o.fr_to_mont_bytes(x) for x in v)
    got = dg.qap_h_from_abc(enc(a), enc(b), enc(c), logn)
    exp = cref.qap_h_from_abc(enc(a), enc(b), enc(c), logn)
    assert np.array_equal(got, exp)
    hv = [o.fr_from_mont_bytes(bytes(got[32 * i:32 * i + 32])) for i in range(n)]
    assert hv[n - 1] == 0                              # deg(h) <= n - 2


def test_fr_spmv_vs_bigint(dg):
    """dg_fr_spmv (the evaluate_constraint map, r1cs_to_qap.rs:150-186) against Python integers: a random CSR matrix with
    empty rows and repeated columns, and the A / B / C matrices of a squaring-chain circuit read by crypto_b200/r1cs.py."""
    import random
    from oracle import bls12_381 as o
    from crypto_b200 import r1cs
    rng = random.Random(99)
    ncols, rows = 37, 50
    w = [1] + [rng.randrange(o.R) for _ in range(ncols - 1)]
    rp, col, co = [0], [], []
    for i in range(rows):
        for _ in range(0 if i % 7 == 3 else rng.randrange(1, 6)):
            col.append(rng.randrange(ncols)); co.append(rng.choice([1, o.R - 1, rng.randrange(o.R)]))
        rp.append(len(col))
    mont = lambda xs: np.frombuffer(b''.join(o.fr_to_mont_bytes(x) for x in xs), np.uint8)
    got = bytes(dg.fr_spmv(rp, col, mont(co), mont(w)))
    exp = [sum(co[k] * w[col[k]] for k in range(rp[i], rp[i + 1])) % o.R for i in range(rows)]
    assert got == b''.join(o.fr_to_mont_bytes(x) for x in exp)
    # a circuit: x_{i+1} = x_i^2, 64 constraints
    n = 64
    cons = [([(2 + i, 1)], [(2 + i, 1)], [(3 + i if i + 1 < n else 1, 1)]) for i in range(n)]
    f = r1cs.R1CSFile.new(r1cs.write_r1cs(1, 0, 1, n + 2, cons))
    xs = [5]
    for _ in range(n):
        xs.append(xs[-1] * xs[-1] % o.R)
    wit = [1, xs[-1]] + xs[:-1]
    a, b, c = f.evaluate(wit)
    for (rpk, colk, valk), expk in zip(f.matrices(), (a, b, c)):
        coeffs = [int.from_bytes(bytes(v), 'little') for v in valk]
        gotk = bytes(dg.fr_spmv(rpk, colk, mont(coeffs), mont(wit)))
        assert gotk == b''.join(o.fr_to_mont_bytes(x) for x in expk)
    assert all((x * y - z) % o.R == 0 for x, y, z in zip(a, b, c))
    with pytest.raises(dg.DockGpuError):
        dg.fr_spmv([0, 1], [ncols + 5], mont([1]), mont(w))                      # column out of range
    assert bytes(dg.fr_spmv([0, 1], [0], mont([7]), mont(w))) == o.fr_to_mont_bytes(7)          # still usable
