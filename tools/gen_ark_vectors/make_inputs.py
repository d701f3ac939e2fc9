"""Writes tests/golden/ark_inputs.json: the seeded inputs tools/gen_ark_vectors (Rust, arkworks) turns into
tests/golden/ark_vectors.json.  Points are given in ark's own compressed encoding so the Rust side needs nothing but
deserialize_compressed.  Run from the repo root:  python tools/gen_ark_vectors/make_inputs.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bls12_381 as o          # noqa: E402
from oracle import cref                    # noqa: E402
import numpy as np                         # noqa: E402


def scalars(n, seed):
    a = np.asarray(cref.random_scalars(n, seed)).reshape(-1, 32)
    return [bytes(r).hex() for r in a]


def g1_points(n, seed):
    aff = bytes(cref.g1_generator_muls(cref.random_scalars(n, seed)))
    return [o.g1_serialize(o.g1_from_bytes(aff[96 * i:96 * i + 96]), True).hex() for i in range(n)]


def g2_points(n, seed):
    aff = bytes(cref.g2_generator_muls(cref.random_scalars(n, seed)))
    return [o.g2_serialize(o.g2_from_bytes(aff[192 * i:192 * i + 192]), True).hex() for i in range(n)]


def main():
    r = o.R
    edge = [(0).to_bytes(32, 'little').hex(), (1).to_bytes(32, 'little').hex(), (r - 1).to_bytes(32, 'little').hex(),
            ((r - 1) // 2).to_bytes(32, 'little').hex(), ((r + 1) // 2).to_bytes(32, 'little').hex()]
    inf1 = o.g1_serialize(None, True).hex()
    pts33 = g1_points(33, 9001)
    data = {
        'msm_g1': [
            {'name': 'n1', 'bases': g1_points(1, 9101), 'scalars': scalars(1, 9102)},
            {'name': 'n33_edge_scalars', 'bases': pts33, 'scalars': edge + scalars(28, 9103)},
            {'name': 'n100_with_identity_and_repeats', 'bases': [inf1] + pts33[:3] * 11 + pts33, 'scalars': scalars(67, 9104)},
            {'name': 'n1000', 'bases': g1_points(1000, 9105), 'scalars': scalars(1000, 9106)},
            {'name': 'n2048', 'bases': g1_points(2048, 9107), 'scalars': scalars(2048, 9108)},
        ],
        'msm_g2': [
            {'name': 'n1', 'bases': g2_points(1, 9201), 'scalars': scalars(1, 9202)},
            {'name': 'n257', 'bases': g2_points(257, 9203), 'scalars': scalars(257, 9204)},
        ],
        'fixed_base_g1': [
            {'name': 'hint10', 'point': g1_points(1, 9301), 'hint': 10, 'scalars': edge + scalars(5, 9302)},
            {'name': 'hint10000', 'point': g1_points(1, 9303), 'hint': 10000, 'scalars': edge + scalars(59, 9304)},
        ],
        'mul_bigint_g1': [
            {'name': 'n16', 'points': g1_points(16, 9401), 'scalars': edge + scalars(11, 9402)},
        ],
        'pairing': [
            {'name': 'pairs1', 'g1': g1_points(1, 9501), 'g2': g2_points(1, 9502)},
            {'name': 'generators', 'g1': [o.g1_serialize(o.G1_GEN, True).hex()], 'g2': [o.g2_serialize(o.G2_GEN, True).hex()]},
            {'name': 'pairs3', 'g1': g1_points(3, 9503), 'g2': g2_points(3, 9504)},
            {'name': 'pairs9', 'g1': g1_points(9, 9505), 'g2': g2_points(9, 9506)},
        ],
    }
    path = os.path.join(ROOT, 'tests', 'golden', 'ark_inputs.json')
    with open(path, 'w') as f:
        json.dump(data, f, indent=0)
    print(path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
