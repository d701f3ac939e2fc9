"""Shared input generators / converters for the parity tests (seeded, reproducible)."""
import numpy as np

from oracle import bls12_381 as o
from oracle import cref

R = o.R


def scalars_bytes(ints):
    return np.frombuffer(b''.join(o.scalar_to_bytes(s) for s in ints), dtype=np.uint8)


def rand_scalars(n, seed):
    """n x 32 B canonical scalars, uniform in [0, r)."""
    return cref.random_scalars(n, seed)


def g1_bases(n, seed):
    """n affine G1 points k_i * G with seeded k_i (known discrete logs), plus the k_i."""
    ks = cref.random_scalars(n, seed)
    return cref.g1_generator_muls(ks), ks


def g2_bases(n, seed):
    ks = cref.random_scalars(n, seed)
    return cref.g2_generator_muls(ks), ks


def ints_of(scalar_bytes):
    a = np.asarray(scalar_bytes, dtype=np.uint8).reshape(-1, 32)
    return [int.from_bytes(bytes(r), 'little') for r in a]


def affine_g1(jac):
    """Canonical affine bytes of Jacobian G1 record(s) via the oracle (the comparison form)."""
    return bytes(cref.normalize_batch_g1(cref._as_np(jac)))


def affine_g2(jac):
    return bytes(cref.normalize_batch_g2(cref._as_np(jac)))


def known_dlog_msm_g1(ks, ss):
    """(sum s_i k_i mod r) * G as affine bytes: O(n) independent check of an MSM over k_i*G."""
    tot = sum(a * b for a, b in zip(ints_of(ks), ints_of(ss))) % R
    return bytes(cref.g1_generator_muls(scalars_bytes([tot])))


def known_dlog_msm_g2(ks, ss):
    tot = sum(a * b for a, b in zip(ints_of(ks), ints_of(ss))) % R
    return bytes(cref.g2_generator_muls(scalars_bytes([tot])))


def neg_g1(aff):
    """-P of an affine G1 record via the big-int oracle."""
    p = o.g1_from_bytes(bytes(aff))
    return o.g1_to_bytes(o.E1.neg(p))
