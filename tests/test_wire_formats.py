"""CPU: the oracle's restatement of the ark-serialize / Zcash point encodings (row f4) is pinned on the
publicly known compressed generators and is self-consistent (round trips, fast subgroup tests ==
the definition [r]P == O)."""
from oracle import bls12_381 as o

G1_GEN_COMPRESSED = ('97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58'
                     '6c55e83ff97a1aeffb3af00adb22c6bb')
G2_GEN_COMPRESSED = ('93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049'
                     '334cf11213945d57e5ac7d055d042b7e024aa2b2f08f0a91260805272dc51051'
                     'c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8')


def test_known_generator_encodings():
    assert o.g1_serialize(o.G1_GEN).hex() == G1_GEN_COMPRESSED
    assert o.g2_serialize(o.G2_GEN).hex() == G2_GEN_COMPRESSED
    assert o.g1_deserialize(bytes.fromhex(G1_GEN_COMPRESSED)) == (o.SER_OK, o.G1_GEN)
    assert o.g2_deserialize(bytes.fromhex(G2_GEN_COMPRESSED)) == (o.SER_OK, o.G2_GEN)
    assert o.g1_serialize(None) == bytes([0xC0]) + bytes(47)
    assert o.g2_serialize(None, compressed=False) == bytes([0x40]) + bytes(191)


def test_round_trips_and_sign_flag():
    for k in (1, 2, 3, 0xDEADBEEF, o.R - 1):
        p, q = o.E1.mul(o.G1_GEN, k), o.E2.mul(o.G2_GEN, k)
        for comp in (True, False):
            assert o.g1_deserialize(o.g1_serialize(p, comp), comp) == (o.SER_OK, p)
            assert o.g2_deserialize(o.g2_serialize(q, comp), comp) == (o.SER_OK, q)
        # P and -P differ exactly in the sort flag
        a, b = o.g1_serialize(p), o.g1_serialize(o.E1.neg(p))
        assert a[1:] == b[1:] and (a[0] ^ b[0]) == 0x20
        a, b = o.g2_serialize(q), o.g2_serialize(o.E2.neg(q))
        assert a[1:] == b[1:] and (a[0] ^ b[0]) == 0x20


def test_fast_subgroup_tests_match_definition():
    for s in range(4):
        p, q = o.curve_point_from_x(False, 900 + s), o.curve_point_from_x(True, 950 + s)
        assert o.E1.on_curve(p) and o.E2.on_curve(q)
        assert o.g1_in_subgroup(p) == o.g1_in_subgroup_fast(p)
        assert o.g2_in_subgroup(q) == o.g2_in_subgroup_fast(q)
        assert o.g1_deserialize(o.g1_serialize(p))[0] == (o.SER_OK if o.g1_in_subgroup(p) else o.SER_NOT_IN_SUBGROUP)
        assert o.g1_deserialize(o.g1_serialize(p), validate=False) == (o.SER_OK, p)
    for k in (5, 77):
        assert o.g1_in_subgroup_fast(o.E1.mul(o.G1_GEN, k)) and o.g2_in_subgroup_fast(o.E2.mul(o.G2_GEN, k))


def test_rejections():
    bad_x = (o.P).to_bytes(48, 'big')
    assert o.g1_deserialize(bytes([bad_x[0] | 0x80]) + bad_x[1:])[0] == o.SER_MALFORMED          # x >= p
    assert o.g1_deserialize(bytes.fromhex(G1_GEN_COMPRESSED), compressed=False)[0] == o.SER_MALFORMED
    assert o.g1_deserialize(bytes([0xC0]) + bytes(46) + b'\x01')[0] == o.SER_MALFORMED            # stray bits
    x = 1
    while o.fp_sqrt((x ** 3 + 4) % o.P) is not None:
        x += 1
    enc = bytearray(x.to_bytes(48, 'big')); enc[0] |= 0x80
    assert o.g1_deserialize(bytes(enc))[0] == o.SER_NOT_ON_CURVE


def test_c_restatement_matches_bigint_oracle(cref):
    import numpy as np
    aff = cref.g1_generator_muls(cref.random_scalars(40, 77))
    pts = [o.g1_from_bytes(bytes(aff[96 * i:96 * i + 96])) for i in range(40)]
    pts += [None, o.E1.neg(pts[0])] + [o.curve_point_from_x(False, 10 + s) for s in range(3)]
    recs = [o.g1_serialize(p) for p in pts]
    recs.append(bytes([0x80 | (o.P >> 376)]) + (o.P).to_bytes(48, 'big')[1:])       # x = p: malformed
    for validate in (True, False):
        out, st = cref.g1_deserialize_compressed(np.frombuffer(b''.join(recs), np.uint8), validate)
        exp = [o.g1_deserialize(r, True, validate) for r in recs]
        assert list(st) == [e[0] for e in exp]
        assert bytes(out) == b''.join(o.g1_to_bytes(e[1]) for e in exp)


def _wire_gold():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'wire_vectors.json')))


def test_oracle_matches_committed_wire_vectors():
    g = _wire_gold()
    assert g['public']['g1_generator_compressed'] == G1_GEN_COMPRESSED and g['public']['g2_generator_compressed'] == G2_GEN_COMPRESSED
    for name, rec, from_b, ser, deser in (('g1', 96, o.g1_from_bytes, o.g1_serialize, o.g1_deserialize),
                                          ('g2', 192, o.g2_from_bytes, o.g2_serialize, o.g2_deserialize)):
        aff = bytes.fromhex(g[name]['affine_montgomery'])
        pts = [from_b(aff[i:i + rec]) for i in range(0, len(aff), rec)]
        assert b''.join(ser(p, True) for p in pts).hex() == g[name]['compressed']
        assert b''.join(ser(p, False) for p in pts).hex() == g[name]['uncompressed']
        bad = bytes.fromhex(g[name]['rejected_compressed'])
        recs = [bad[i:i + rec // 2] for i in range(0, len(bad), rec // 2)]
        assert [deser(r, True, True)[0] for r in recs] == g[name]['rejected_status_validated']
        assert [deser(r, True, False)[0] for r in recs] == g[name]['rejected_status_unvalidated']


def test_c_restatement_g2_matches_bigint_oracle(cref):
    import numpy as np
    aff = cref.g2_generator_muls(cref.random_scalars(16, 78))
    pts = [o.g2_from_bytes(bytes(aff[192 * i:192 * i + 192])) for i in range(16)]
    pts += [None, o.E2.neg(pts[0])] + [o.curve_point_from_x(True, 20 + s) for s in range(3)]
    recs = [o.g2_serialize(p) for p in pts]
    g = _wire_gold()['g2']
    bad = bytes.fromhex(g['rejected_compressed'])
    recs += [bad[i:i + 96] for i in range(0, len(bad), 96)]
    for validate in (True, False):
        out, st = cref.g2_deserialize_compressed(np.frombuffer(b''.join(recs), np.uint8), validate)
        exp = [o.g2_deserialize(r, True, validate) for r in recs]
        assert list(st) == [e[0] for e in exp]
        assert bytes(out) == b''.join(o.g2_to_bytes(e[1]) for e in exp)
