"""Host-side mirror of the reference's pairing-facing interface over the C ABI.

  * ark_ec::pairing::Pairing::{multi_miller_loop, final_exponentiation, multi_pairing, pairing}
  * utils::randomized_pairing_check::RandomizedPairingChecker  (utils/src/randomized_pairing_check.rs:24-215)

GT elements / Miller-loop outputs are 576-byte Fp12 records (ark field order, Montgomery).
PairingOutput is written additively in arkworks (zero() == Fp12::one, `+=` is multiplication,
mul_bigint is exponentiation); the same operations are used here.
"""
import numpy as np

from . import lib
from .msm import R_MODULUS, points_to_bytes, normalize_batch, mul_bigint_batch, G1


def multi_miller_loop(a, b):
    """Bls12::multi_miller_loop: pairs with an identity on either side are skipped."""
    return bytes(lib.multi_miller_loop(points_to_bytes(a), points_to_bytes(b)))


def final_exponentiation(f):
    """None iff f == 0 (ark returns Option)."""
    r = lib.final_exponentiation(f)
    return None if r is None else bytes(r)


def multi_pairing(a, b):
    return bytes(lib.multi_pairing(points_to_bytes(a), points_to_bytes(b)))


def pairing(p, q):
    return multi_pairing([p], [q])


_GT_ONE = None


def gt_zero():
    """PairingOutput::zero() = the multiplicative identity of Fp12."""
    global _GT_ONE
    if _GT_ONE is None:
        _GT_ONE = bytes(lib.multi_pairing(b'', b''))
    return _GT_ONE


def gt_add(x, y):
    """PairingOutput + PairingOutput (Fp12 multiplication)."""
    return bytes(lib.fp12_mul(x, y))


def gt_mul_bigint(x, m):
    """PairingOutput::mul_bigint."""
    return bytes(lib.gt_pow(x, int(m % R_MODULUS).to_bytes(32, 'little')))


class RandomizedPairingChecker:
    """Random-linear-combination of pairing equations with ONE final exponentiation.
    Field-for-field the reference struct: left (Miller-loop accumulator), right (GT target),
    lazy, pending, random, current_random."""

    def __init__(self, random, lazy):
        self.left = gt_zero()                 # MillerLoopOutput(Fp12::one)
        self.right = gt_zero()                # PairingOutput::zero()
        self.lazy = lazy
        self.pending = ([], [])
        self.random = random % R_MODULUS
        self.current_random = 1

    @classmethod
    def new(cls, random, lazy):
        return cls(random, lazy)

    # -- helpers ------------------------------------------------------------------------------
    def _scaled(self, pts, m):
        """[a_i * m] as affine records (a.mul_bigint(m) then G1Prepared::from)."""
        pts = points_to_bytes(pts)
        k = pts.size // 96
        if k == 0:
            return []
        sc = np.frombuffer(int(m % R_MODULUS).to_bytes(32, 'little') * k, dtype=np.uint8)
        aff = normalize_batch(mul_bigint_batch(pts, sc), G1)
        return [bytes(aff[96 * i:96 * i + 96]) for i in range(k)]

    def _advance(self):
        self.current_random = self.current_random * self.random % R_MODULUS

    # -- reference API ------------------------------------------------------------------------
    def add_sources_and_target(self, a, b, out):
        self.add_multiple_sources_and_target([a], [b], out)

    def add_multiple_sources_and_target(self, a, b, out):
        self.add_multiple_sources_and_target_with_laziness_choice(a, b, out, self.lazy)

    def add_multiple_sources(self, a, b, c, d):
        self.add_multiple_sources_with_laziness_choice(a, b, c, d, self.lazy)

    def add_sources(self, a, b, c, d):
        self.add_sources_with_laziness_choice(a, b, c, d, self.lazy)

    def add_multiple_sources_and_target_with_laziness_choice(self, a, b, out, lazy):
        m = self.current_random
        a_m = self._scaled(a, m)
        b = [bytes(x) for x in b]
        if lazy:
            self.pending[0].extend(a_m)
            self.pending[1].extend(b)
        else:
            self.left = gt_add(self.left, multi_miller_loop(a_m, b))
        self.right = gt_add(self.right, gt_mul_bigint(out, m))
        self._advance()

    def add_multiple_sources_with_laziness_choice(self, a, b, c, d, lazy):
        m = self.current_random
        a_m = self._scaled(a, m)
        c_m = self._scaled(c, -m)            # -c.mul_bigint(m) == c * (r - m)
        b = [bytes(x) for x in b]
        d = [bytes(x) for x in d]
        if lazy:
            self.pending[0].extend(a_m); self.pending[1].extend(b)
            self.pending[0].extend(c_m); self.pending[1].extend(d)
        else:
            self.left = gt_add(self.left, multi_miller_loop(a_m, b))
            self.left = gt_add(self.left, multi_miller_loop(c_m, d))
        self._advance()

    def add_sources_with_laziness_choice(self, a, b, c, d, lazy):
        m = self.current_random
        am = self._scaled([a], m)
        cm = self._scaled([c], -m)
        if lazy:
            self.pending[0].extend(am + cm)
            self.pending[1].extend([bytes(b), bytes(d)])
        else:
            self.left = gt_add(self.left, multi_miller_loop(am + cm, [bytes(b), bytes(d)]))
        self._advance()

    def verify(self):
        assert len(self.pending[0]) == len(self.pending[1])
        left = self.left
        if self.pending[0]:
            p = multi_miller_loop(self.pending[0], self.pending[1])
            left = gt_add(p, self.left)
        fe = final_exponentiation(left)
        assert fe is not None                 # reference unwraps
        return fe == self.right
