"""GPU parity of the wire-format kernels (csrc/serialize.cu, row f4) against the oracle: encodings,
decodings, every rejection class, and a round trip at 2^16 points."""
import numpy as np
import pytest

from oracle import bls12_381 as o
from tests import helpers as h
from tests.test_wire_formats import G1_GEN_COMPRESSED, G2_GEN_COMPRESSED

pytestmark = pytest.mark.gpu


def _pts(aff, g2):
    rec = 192 if g2 else 96
    f = o.g2_from_bytes if g2 else o.g1_from_bytes
    a = bytes(aff)
    return [f(a[i:i + rec]) for i in range(0, len(a), rec)]


@pytest.mark.parametrize('g2', [False, True])
@pytest.mark.parametrize('compressed', [True, False])
def test_serialize_vs_oracle(dg, g2, compressed):
    n = 40
    aff, _ = (h.g2_bases if g2 else h.g1_bases)(n, 811 + g2)
    aff = aff.copy()
    rec = 192 if g2 else 96
    aff[3 * rec:4 * rec] = 0                                   # an identity in the middle
    ser = o.g2_serialize if g2 else o.g1_serialize
    exp = b''.join(ser(p, compressed) for p in _pts(aff, g2))
    assert bytes(dg.serialize_points(aff, g2=g2, compressed=compressed)) == exp


def test_known_generator_encodings(dg):
    g1 = o.g1_to_bytes(o.G1_GEN)
    g2 = o.g2_to_bytes(o.G2_GEN)
    assert bytes(dg.serialize_points(g1)).hex() == G1_GEN_COMPRESSED
    assert bytes(dg.serialize_points(g2, g2=True)).hex() == G2_GEN_COMPRESSED
    out, st, bad = dg.deserialize_points(bytes.fromhex(G1_GEN_COMPRESSED))
    assert bytes(out) == g1 and list(st) == [0] and bad == 0
    out, st, bad = dg.deserialize_points(bytes.fromhex(G2_GEN_COMPRESSED), g2=True)
    assert bytes(out) == g2 and list(st) == [0] and bad == 0


@pytest.mark.parametrize('g2', [False, True])
@pytest.mark.parametrize('compressed', [True, False])
def test_deserialize_vs_oracle_incl_rejections(dg, g2, compressed):
    ser, deser = (o.g2_serialize, o.g2_deserialize) if g2 else (o.g1_serialize, o.g1_deserialize)
    to_bytes = o.g2_to_bytes if g2 else o.g1_to_bytes
    aff, _ = (h.g2_bases if g2 else h.g1_bases)(24, 821 + g2)
    pts = _pts(aff, g2) + [None]
    pts += [(o.E2 if g2 else o.E1).neg(p) for p in pts[:6]]                      # both sort-flag values
    pts += [o.curve_point_from_x(g2, 700 + s) for s in range(5)]                 # on the curve, outside the subgroup
    recs = [ser(p, compressed) for p in pts]
    csz = 96 if g2 else 48
    big = (o.P + 5).to_bytes(48, 'big')
    first = bytearray(recs[0])
    bad1 = bytearray(first); bad1[:48] = big; bad1[0] |= (0x80 if compressed else 0)       # coordinate >= p
    bad2 = bytearray(first); bad2[0] ^= 0x80                                                # wrong compression flag
    bad3 = bytearray(len(first)); bad3[0] = 0xC0 if compressed else 0x40; bad3[-1] = 1      # infinity with stray bits
    recs += [bytes(bad1), bytes(bad2), bytes(bad3)]
    if compressed:                                                                           # an x without a root
        x = 1
        while True:
            xx = (x, 0) if g2 else x
            root = o.fp2_sqrt(o.fp2_add(o.fp2_mul(o.fp2_sqr(xx), xx), (4, 4))) if g2 else o.fp_sqrt((x ** 3 + 4) % o.P)
            if root is None:
                break
            x += 1
        enc = bytearray((bytes(48) if g2 else b'') + x.to_bytes(48, 'big')); enc[0] |= 0x80
        recs.append(bytes(enc))
    else:                                                                                    # y tampered: not on the curve
        t = bytearray(first); t[-1] ^= 1
        recs.append(bytes(t))
    assert all(len(r) == (csz if compressed else 2 * csz) for r in recs)
    for validate in (True, False):
        out, st, bad = dg.deserialize_points(b''.join(recs), g2=g2, compressed=compressed, validate=validate)
        exp = [deser(r, compressed, validate) for r in recs]
        assert list(st) == [e[0] for e in exp]
        assert bad == sum(1 for e in exp if e[0] != 0)
        assert bytes(out) == b''.join(to_bytes(e[1]) for e in exp)
    assert {e[0] for e in exp} == {0, 1, 2}                    # validate=False run: no subgroup rejections
    st_v = dg.deserialize_points(b''.join(recs), g2=g2, compressed=compressed, validate=True)[1]
    assert set(st_v) == {0, 1, 2, 3}


@pytest.mark.parametrize('g2', [False, True])
def test_round_trip_large(dg, g2):
    n = 1 << (13 if g2 else 16)
    aff, _ = (h.g2_bases if g2 else h.g1_bases)(n, 831 + g2)
    for compressed in (True, False):
        enc = dg.serialize_points(aff, g2=g2, compressed=compressed)
        out, st, bad = dg.deserialize_points(enc, g2=g2, compressed=compressed, validate=True)
        assert bad == 0 and not st.any()
        assert np.array_equal(out, np.asarray(aff).reshape(-1))
