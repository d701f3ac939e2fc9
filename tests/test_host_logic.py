"""CPU: host-side mirror logic that needs no GPU (window rules, marshalling, checker bookkeeping)."""
import numpy as np
import pytest

from crypto_b200 import msm
from oracle import bls12_381 as o


def test_window_size_rule_matches_ark():
    for n in (1, 10, 31, 32, 33, 1000, 10000, 1 << 19):
        assert msm.WindowTable.window_size_for(n) == o.fixed_base_window_size(n)


def test_scalar_marshalling():
    b = msm.scalars_to_bytes([0, 1, o.R - 1, o.R + 5])
    assert b.size == 128
    assert bytes(b[32:64]) == (1).to_bytes(32, 'little')
    assert bytes(b[96:128]) == (5).to_bytes(32, 'little')       # reduced mod r like Fr
    raw = np.arange(64, dtype=np.uint8)
    assert msm.scalars_to_bytes(raw) is not None and msm.scalars_to_bytes(bytes(raw)).size == 64


def test_mult_checker_dedups_by_x_coordinate():
    """utils/src/randomized_mult_checker.rs:107-125: P and -P share one entry."""
    p = o.E1.mul(o.G1_GEN, 5)
    pb, nb = o.g1_to_bytes(p), o.g1_to_bytes(o.E1.neg(p))
    ck = msm.RandomizedMultChecker(7)
    ck._add(pb, 10)
    ck._add(pb, 5)
    ck._add(nb, 3)
    ck._add(bytes(96), 99)           # identity ignored
    assert len(ck) == 1
    (scalar, point), = ck.args.values()
    assert scalar == 12 and point == pb
    ck.add_1(pb, 2, o.g1_to_bytes(o.E1.mul(p, 2)))
    assert ck.current_random == 7


def test_length_mismatch_error():
    v = msm.VariableBaseMSM()
    with pytest.raises(msm.LengthMismatch) as ei:
        v.msm(bytes(96 * 3), [1, 2])
    assert ei.value.min_len == 2


def test_serde_framing_errors_without_gpu():
    """Length framing is checked before anything reaches the GPU."""
    import struct
    import pytest
    from crypto_b200 import serde
    with pytest.raises(serde.SerializationError) as ei:
        serde.deserialize_vec(b'\x01\x02', serde.G1)
    assert ei.value.kind == 'IoError'
    with pytest.raises(serde.SerializationError) as ei:
        serde.deserialize_vec(struct.pack('<Q', 3) + bytes(48 * 2), serde.G1)
    assert ei.value.kind == 'IoError'
    assert serde.serialized_size(5, serde.G1) == 8 + 5 * 48 and serde.serialized_size(5, serde.G2, compressed=False) == 8 + 5 * 192
    assert len(serde.deserialize_vec(struct.pack('<Q', 0), serde.G2)) == 0


def test_msm_digit_recoding_spec():
    """The digit count the kernels use (msm_ndigits) is enough for EVERY canonical scalar at every window size, and
    the recoding reproduces the scalar: sum d_j 2^(c j) == min(s, r - s), with the sign folded into `flip`."""
    import random
    from crypto_b200 import msm
    r = msm.R_MODULUS
    rng = random.Random(7)
    edge = [0, 1, 2, (r - 1) // 2 - 1, (r - 1) // 2, (r - 1) // 2 + 1, r - 2, r - 1, (1 << 254) - 1, 1 << 253, (1 << 254) % r]
    for c in range(2, 25):
        nd = msm.msm_ndigits(c)
        assert nd * c >= 254 and (nd - 1) * c <= 254
        half = 1 << (c - 1)
        worst = [sum(half << (c * j) for j in range(nd)) % r]            # every raw digit at the carry threshold
        for s in edge + worst + [rng.randrange(r) for _ in range(200)]:
            flip, d = msm.msm_recode(s % r, c)
            assert len(d) == nd and all(-half < x <= half for x in d)
            assert sum(x << (c * j) for j, x in enumerate(d)) == (r - s % r if flip else s % r)
            assert (r - s % r if flip else s % r) <= (r - 1) // 2 or s % r == 0
    import pytest
    with pytest.raises(ValueError):
        msm.msm_recode(r, 16)


def test_glv_split_is_exact_and_short():
    """The GLV mirror of k_digits: s == sign1 k1 + sign2 k2 lambda (mod r), both halves below 2^127, and their signed
    digits never carry out of glv_ndigits(c) windows."""
    import random
    from crypto_b200 import msm
    r, x2, lam = msm.R_MODULUS, msm.GLV_X2, msm.GLV_LAMBDA
    assert (lam * lam + lam + 1) % r == 0
    rng = random.Random(11)
    cases = [0, 1, r - 1, (r - 1) // 2, (r + 1) // 2, x2, x2 - 1, x2 + 1, x2 // 2, x2 // 2 + 1, lam, lam + 1, r - lam, 3 * x2 + x2 // 2]
    cases += [rng.randrange(r) for _ in range(3000)] + [rng.randrange(1 << k) for k in range(1, 255, 3)]
    for s in cases:
        s1, k1, s2, k2 = msm.glv_split(s)
        assert (s1 * k1 + s2 * k2 * lam) % r == s and k1 < (1 << 127) and k2 < (1 << 127)
        for c in (4, 11, 13, 16, 17, 20, 23):
            for k in (k1, k2):
                carry, half = 0, 1 << (c - 1)
                for j in range(msm.glv_ndigits(c)):
                    d = ((k >> (c * j)) & ((1 << c) - 1)) + carry
                    carry = 1 if d > half else 0
                assert carry == 0
    with pytest.raises(ValueError):
        msm.glv_split(r)
