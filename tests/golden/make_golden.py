#!/usr/bin/env python3
"""Generates tests/golden/vectors.json from the big-integer oracle (oracle/bls12_381.py).

The reference (docknetwork/crypto) has no golden vectors for this path and cannot be executed
here (no Rust toolchain; arkworks not vendored), so these vectors are produced by the
independent Python big-int implementation: naive double-and-add MSM, the textbook pairing
(polynomial-ring Fp12, affine Miller loop, plain pow) cubed to the arkworks convention, and the
public generator constants.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import bls12_381 as o  # noqa: E402


def poly_to_fp12(poly):
    """Inverse of o.fp12_to_poly: polynomial in w (deg < 12) -> tower element."""
    # coefficient of w^k (k<6) = x - y, of w^(k+6) = y for a_k = x + y u
    tower = [[None] * 3 for _ in range(2)]
    for k in range(6):
        y = poly[k + 6] % o.P
        x = (poly[k] + y) % o.P
        tower[k & 1][k >> 1] = (x, y)
    return (tuple(tower[0]), tuple(tower[1]))


def main():
    rng = o.SplitMix64(0xD0C4C0DE)
    v = {}
    v['g1_generator_compressed'] = o.g1_compressed(o.G1_GEN).hex()
    v['g1_generator'] = o.g1_to_bytes(o.G1_GEN).hex()
    v['g2_generator'] = o.g2_to_bytes(o.G2_GEN).hex()
    # MSM G1: 12 terms incl. zero scalar, scalar one, r-1, identity base, repeated base, P and -P
    ks = [rng.scalar() for _ in range(12)]
    pts = [o.E1.mul(o.G1_GEN, k) for k in ks]
    pts[3] = None
    pts[5] = pts[4]
    pts[7] = o.E1.neg(pts[6])
    sc = [rng.scalar() for _ in range(12)]
    sc[0], sc[1], sc[2] = 0, 1, o.R - 1
    sc[7] = sc[6]
    v['msm_g1'] = {
        'bases': b''.join(o.g1_to_bytes(p) for p in pts).hex(),
        'scalars': b''.join(o.scalar_to_bytes(s) for s in sc).hex(),
        'result_affine': o.g1_to_bytes(o.E1.msm_naive(pts, sc)).hex(),
    }
    ks2 = [rng.scalar() for _ in range(6)]
    pts2 = [o.E2.mul(o.G2_GEN, k) for k in ks2]
    pts2[2] = None
    sc2 = [rng.scalar() for _ in range(6)]
    sc2[0] = 0
    v['msm_g2'] = {
        'bases': b''.join(o.g2_to_bytes(p) for p in pts2).hex(),
        'scalars': b''.join(o.scalar_to_bytes(s) for s in sc2).hex(),
        'result_affine': o.g2_to_bytes(o.E2.msm_naive(pts2, sc2)).hex(),
    }
    # fixed base: 5 scalars times one point
    base = o.E1.mul(o.G1_GEN, rng.scalar())
    fs = [rng.scalar() for _ in range(5)] + [0, 1]
    v['fixed_base_g1'] = {
        'point': o.g1_to_bytes(base).hex(),
        'scalars': b''.join(o.scalar_to_bytes(s) for s in fs).hex(),
        'results_affine': b''.join(o.g1_to_bytes(o.E1.mul(base, s)) for s in fs).hex(),
    }
    # pairing: textbook derivation cubed (arkworks convention), e(aG1, bG2) and a 2-pair product
    a, b = rng.scalar(), rng.scalar()
    p1, q1 = o.E1.mul(o.G1_GEN, a), o.E2.mul(o.G2_GEN, b)
    e_gen = o.pairing_textbook(o.G1_GEN, o.G2_GEN)
    e_ab = o.pairing_textbook(p1, q1)
    assert e_ab == o._poly_pow(e_gen, (a * b) % o.R)
    v['pairing'] = {
        'e_g1_g2': o.fp12_to_bytes(poly_to_fp12(e_gen)).hex(),
        'p': o.g1_to_bytes(p1).hex(), 'q': o.g2_to_bytes(q1).hex(),
        'e_p_q': o.fp12_to_bytes(poly_to_fp12(e_ab)).hex(),
        # e(P, Q) * e(-P, Q) == 1
        'neg_p': o.g1_to_bytes(o.E1.neg(p1)).hex(),
    }
    with open(os.path.join(HERE, 'vectors.json'), 'w') as f:
        json.dump(v, f, indent=1)
    print('wrote vectors.json')


if __name__ == '__main__':
    main()
