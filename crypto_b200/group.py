"""Small group / field helpers for the host-side mirrors (groth16.py, protocols.py).

Everything that costs curve arithmetic goes to the GPU through the C ABI (lib.batch_mul, lib.fold, lib.msm,
lib.normalize_batch); the host only re-packs records, negates a coordinate (p - y on the Montgomery residue, what
arkworks' `-P` does) and does Fr bookkeeping with Python integers -- the work the Rust glue keeps on the CPU.
"""
import numpy as np

from . import lib

P_MODULUS = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FP_ONE_MONT = ((1 << 384) % P_MODULUS).to_bytes(48, 'little')        # ark-ff Fp::one() in memory
FR_R = (1 << 256) % R_MODULUS
FR_R_INV = pow(FR_R, -1, R_MODULUS)


class G1:
    g2 = False
    AFF, PROJ, FIELD = lib.G1_AFF, lib.G1_JAC, 48


class G2:
    g2 = True
    AFF, PROJ, FIELD = lib.G2_AFF, lib.G2_JAC, 96


def fr_to_mont(values):
    """ints -> n x 32 B Fr Montgomery records (the in-memory form of ark-ff's Fr)."""
    return np.frombuffer(b''.join(((int(v) % R_MODULUS) * FR_R % R_MODULUS).to_bytes(32, 'little') for v in values), dtype=np.uint8)


def fr_from_mont(data):
    a = np.asarray(data, dtype=np.uint8).reshape(-1, 32)
    return [int.from_bytes(bytes(r), 'little') * FR_R_INV % R_MODULUS for r in a]


def fr_to_bytes(values):
    """ints -> n x 32 B canonical little-endian (Fr::into_bigint)."""
    return np.frombuffer(b''.join((int(v) % R_MODULUS).to_bytes(32, 'little') for v in values), dtype=np.uint8)


def is_identity(aff):
    return not any(bytes(aff))


def neg(aff, group=G1):
    """-P for one affine record: y -> p - y on every Fp component (Montgomery residues negate like integers mod p)."""
    aff = bytes(aff)
    if is_identity(aff):
        return aff
    half = group.AFF // 2
    out = bytearray(aff[:half])
    for k in range(half // 48):
        y = int.from_bytes(aff[half + 48 * k:half + 48 * k + 48], 'little')
        out += ((P_MODULUS - y) % P_MODULUS).to_bytes(48, 'little')
    return bytes(out)


def to_projective(aff, group=G1):
    """Affine -> ark Projective record (x, y, 1); the identity becomes (1, 1, 0)."""
    aff = bytes(aff)
    one = FP_ONE_MONT + (bytes(48) if group.g2 else b'')
    if is_identity(aff):
        return one + one + bytes(group.FIELD)
    return aff + one


def add(points, group=G1):
    """Sum of projective records (bytes each) on the GPU -> one projective record."""
    return bytes(lib.fold(np.frombuffer(b''.join(bytes(p) for p in points), dtype=np.uint8), g2=group.g2))


def add_affine(points, group=G1):
    return add([to_projective(p, group) for p in points], group)


def into_affine(proj, group=G1):
    return bytes(lib.normalize_batch(np.frombuffer(bytes(proj), dtype=np.uint8), g2=group.g2))


def mul(aff, scalar, group=G1):
    """AffineRepr::mul_bigint for one point -> projective record."""
    return bytes(lib.batch_mul(np.frombuffer(bytes(aff), dtype=np.uint8), fr_to_bytes([scalar]), g2=group.g2))


def mul_affine(aff, scalar, group=G1):
    return into_affine(mul(aff, scalar, group), group)


def split(records, size):
    b = bytes(records)
    return [b[i:i + size] for i in range(0, len(b), size)]
