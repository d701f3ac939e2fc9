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
