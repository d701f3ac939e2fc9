"""Shared by tests/test_ark_vectors.py (oracle, CPU) and tests/test_gpu_ark_vectors.py (GPU): decoding of
tests/golden/ark_inputs.json / ark_vectors.json (tools/gen_ark_vectors).  ark_vectors.json holds arkworks' OWN outputs;
it can only be produced where a Rust toolchain exists, so the tests skip while it is absent (parity stays "unpinned by
the reference") and turn into the reference pin the moment someone commits it."""
import json
import os

import numpy as np
import pytest

from oracle import bls12_381 as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
INPUTS = os.path.join(GOLDEN, 'ark_inputs.json')
VECTORS = os.environ.get('ARK_VECTORS', os.path.join(GOLDEN, 'ark_vectors.json'))


def load():
    if not os.path.exists(VECTORS):
        pytest.skip('tests/golden/ark_vectors.json not generated yet (needs cargo: tools/gen_ark_vectors)')
    return json.load(open(INPUTS)), json.load(open(VECTORS))


def case(section, name):
    return next(c for c in section if c['name'] == name)


def g1_records(hexes):
    """compressed hex -> concatenated Montgomery affine records (the C ABI's input form), via the big-int oracle."""
    return np.frombuffer(b''.join(o.g1_to_bytes(o.g1_deserialize(bytes.fromhex(x), True, False)[1]) for x in hexes), dtype=np.uint8)


def g2_records(hexes):
    return np.frombuffer(b''.join(o.g2_to_bytes(o.g2_deserialize(bytes.fromhex(x), True, False)[1]) for x in hexes), dtype=np.uint8)


def scalar_records(hexes):
    return np.frombuffer(b''.join(bytes.fromhex(x) for x in hexes), dtype=np.uint8)


def g1_hex(aff_record):
    """Montgomery affine record -> ark compressed hex."""
    return o.g1_serialize(o.g1_from_bytes(bytes(aff_record)), True).hex()


def g2_hex(aff_record):
    return o.g2_serialize(o.g2_from_bytes(bytes(aff_record)), True).hex()


def fp12_hex(mont_record):
    """576-byte Montgomery Fp12 record (ark field order) -> ark's serialize_compressed hex (canonical little-endian)."""
    b = bytes(mont_record)
    rinv = pow(1 << 384, -1, o.P)
    return b''.join((int.from_bytes(b[48 * i:48 * i + 48], 'little') * rinv % o.P).to_bytes(48, 'little') for i in range(12)).hex()
