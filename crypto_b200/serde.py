"""Host-side mirror of the reference's (de)serialisation interface for vectors of group elements.

Same names, argument meaning and error behaviour as ark-serialize 0.4's
  CanonicalSerialize::{serialize_compressed, serialize_uncompressed, serialized_size}
  CanonicalDeserialize::{deserialize_compressed, deserialize_uncompressed,
                         deserialize_compressed_unchecked, deserialize_uncompressed_unchecked}
for `Vec<G1Affine>` / `Vec<G2Affine>` and single points, as the reference reaches them through
`utils::serde_utils::ArkObjectBytes` (utils/src/serde_utils.rs:13-33) and the derives on its key types
(legogroth16/src/data_structures.rs:7-189).  A `Vec<T>` is its length as a little-endian u64 followed by
the elements; a point is the Zcash / IETF record (include/dockgpu.h).  The curve work (square roots,
subgroup checks, Montgomery conversion) happens on the GPU through dg_g1/g2_serialize / _deserialize;
this module only frames bytes.  Errors mirror `SerializationError`: `InvalidData` when an element is
malformed, off the curve or outside the subgroup, `UnexpectedEof` / `NotEnoughSpace`-style length errors as
`IoError`.
"""
import struct

import numpy as np

from . import lib
from .msm import G1, G2, points_to_bytes


class SerializationError(Exception):
    def __init__(self, kind, detail=''):
        super().__init__(kind + (': ' + detail if detail else ''))
        self.kind = kind


def _rec(group, compressed):
    return group.AFF // 2 if compressed else group.AFF


def serialized_size(n, group=G1, compressed=True):
    """CanonicalSerialize::serialized_size of a Vec of n points."""
    return 8 + n * _rec(group, compressed)


def serialize_vec(points, group=G1, compressed=True):
    """Vec<Affine>::serialize_compressed / serialize_uncompressed -> bytes."""
    p = points_to_bytes(points)
    if p.size % group.AFF:
        raise SerializationError('InvalidData', 'affine records have the wrong length')
    n = p.size // group.AFF
    body = lib.serialize_points(p, g2=group.g2, compressed=compressed) if n else np.zeros(0, np.uint8)
    return struct.pack('<Q', n) + bytes(body)


def deserialize_vec(data, group=G1, compressed=True, validate=True):
    """Vec<Affine>::deserialize_compressed (validate=True: Validate::Yes) / _unchecked (validate=False) and the
    uncompressed variants -> affine records (numpy uint8, n x 96 / 192 B)."""
    data = bytes(data)
    if len(data) < 8:
        raise SerializationError('IoError', 'unexpected end of input while reading the length')
    (n,) = struct.unpack('<Q', data[:8])
    rec = _rec(group, compressed)
    if len(data) - 8 < n * rec:
        raise SerializationError('IoError', 'unexpected end of input: %d elements announced, %d bytes present' % (n, len(data) - 8))
    if n == 0:
        return np.zeros(0, np.uint8)
    out, status, bad = lib.deserialize_points(data[8:8 + n * rec], g2=group.g2, compressed=compressed, validate=validate)
    if bad:
        first = int(np.flatnonzero(status)[0])
        why = {1: 'malformed encoding', 2: 'not on the curve', 3: 'not in the prime-order subgroup'}[int(status[first])]
        raise SerializationError('InvalidData', 'element %d: %s' % (first, why))
    return out


def serialize_point(point, group=G1, compressed=True):
    """Affine::serialize_compressed / _uncompressed of one point (what ArkObjectBytes writes for a point field)."""
    return bytes(lib.serialize_points(points_to_bytes([point]), g2=group.g2, compressed=compressed))


def deserialize_point(data, group=G1, compressed=True, validate=True):
    data = bytes(data)
    if len(data) < _rec(group, compressed):
        raise SerializationError('IoError', 'unexpected end of input')
    out, status, bad = lib.deserialize_points(data[:_rec(group, compressed)], g2=group.g2, compressed=compressed, validate=validate)
    if bad:
        raise SerializationError('InvalidData', {1: 'malformed encoding', 2: 'not on the curve', 3: 'not in the prime-order subgroup'}[int(status[0])])
    return out


__all__ = ['G1', 'G2', 'SerializationError', 'serialized_size', 'serialize_vec', 'deserialize_vec', 'serialize_point',
           'deserialize_point']
