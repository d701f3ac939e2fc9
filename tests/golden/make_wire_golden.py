#!/usr/bin/env python3
"""Generates tests/golden/wire_vectors.json (row f4: ark-serialize / Zcash encodings of BLS12-381 points).

Two kinds of entries:
  * `public`: the compressed generators as published with the curve (external pins, typed in, not computed);
  * everything else: produced by the big-integer oracle (oracle/bls12_381.py), whose encoder reproduces the
    public entries -- multiples of the generators in all four encodings, the identity, points on the curve
    outside the prime-order subgroup and malformed records with the status a validating decoder must return
    (0 ok, 1 malformed, 2 not on the curve, 3 not in the subgroup).
Run:  python tests/golden/make_wire_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import bls12_381 as o  # noqa: E402

PUBLIC = {
    'g1_generator_compressed': '97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb',
    'g2_generator_compressed': '93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e'
                               '024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8',
}


def main():
    assert o.g1_serialize(o.G1_GEN).hex() == PUBLIC['g1_generator_compressed']
    assert o.g2_serialize(o.G2_GEN).hex() == PUBLIC['g2_generator_compressed']
    v = {'public': PUBLIC}
    ks = [1, 2, 3, 0xDEADBEEF, o.R - 1, o.R - 2, 0x1234567890ABCDEF1234567890ABCDEF]
    for name, E, G, ser, to_b in (('g1', o.E1, o.G1_GEN, o.g1_serialize, o.g1_to_bytes), ('g2', o.E2, o.G2_GEN, o.g2_serialize, o.g2_to_bytes)):
        pts = [E.mul(G, k) for k in ks] + [None]
        v[name] = {
            'affine_montgomery': b''.join(to_b(p) for p in pts).hex(),
            'compressed': b''.join(ser(p, True) for p in pts).hex(),
            'uncompressed': b''.join(ser(p, False) for p in pts).hex(),
        }
        g2 = name == 'g2'
        deser = o.g2_deserialize if g2 else o.g1_deserialize
        off = [o.curve_point_from_x(g2, 4000 + s) for s in range(3)]
        good = ser(pts[0], True)
        bad = [ser(p, True) for p in off]
        t = bytearray(good); t[0] ^= 0x80; bad.append(bytes(t))                        # compression flag cleared
        t = bytearray(len(good)); t[0] = 0xC0; t[-1] = 1; bad.append(bytes(t))         # infinity with stray bits
        t = bytearray(good); t[:48] = (o.P + 1).to_bytes(48, 'big'); t[0] |= 0x80; bad.append(bytes(t))   # coordinate >= p
        x = 1
        while True:                                                                     # an x without a point
            xx = (x, 0) if g2 else x
            rhs = o.fp2_add(o.fp2_mul(o.fp2_sqr(xx), xx), (4, 4)) if g2 else (x ** 3 + 4) % o.P
            if (o.fp2_sqrt(rhs) if g2 else o.fp_sqrt(rhs)) is None:
                break
            x += 1
        t = bytearray((bytes(48) if g2 else b'') + x.to_bytes(48, 'big')); t[0] |= 0x80; bad.append(bytes(t))
        v[name]['rejected_compressed'] = b''.join(bad).hex()
        v[name]['rejected_status_validated'] = [deser(r, True, True)[0] for r in bad]
        v[name]['rejected_status_unvalidated'] = [deser(r, True, False)[0] for r in bad]
    with open(os.path.join(HERE, 'wire_vectors.json'), 'w') as f:
        json.dump(v, f, indent=1)
    print('wrote wire_vectors.json')


if __name__ == '__main__':
    main()
