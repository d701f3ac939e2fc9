"""Host-side mirror of the protocol-level callers of the hot path, over the C ABI.

Same names, argument meaning and relations as
  * schnorr_pok/src/pok_generalized_pedersen.rs:83-161   SchnorrCommitment::{new, response}, SchnorrResponse::is_valid
  * schnorr_pok/src/discrete_log.rs:84-254               PokPedersenCommitmentProtocol / PokPedersenCommitment
  * bbs_plus/src/setup.rs:128-163                        SignatureParamsG1::{commit_to_messages, b}
  * bbs_plus/src/signature.rs:138-211, 272-296           SignatureG1::{new, verify}
  * bbs_plus/src/proof.rs:159-251, 478-611               PoKOfSignatureG1Protocol::{init, gen_proof},
                                                         PoKOfSignatureG1Proof::{verify, verify_with_randomized_pairing_checker}
  * vb_accumulator/src/batch_utils.rs:81-470, 498-736    Poly_d / Poly_v_A / Poly_v_D / Poly_v_AD, Omega::{new, evaluate}
  * vb_accumulator/src/witness.rs:165-284, 290-345       batch witness updates with the secret key / with Omega
  * vb_accumulator/src/positive.rs:401-425               verify_membership_given_accumulated
Every MSM, scalar multiplication, fixed-base multiplication, normalisation and pairing runs on the GPU; the host keeps
the Fr arithmetic (Python integers), exactly the split of the Rust glue.  Randomness is passed in explicitly so the
parity tests are reproducible.
"""
import hashlib
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import group as gp
from . import lib
from . import msm as msm_mod
from .group import G1, G2, R_MODULUS


def _msm_unchecked(bases, scalars, group=G1):
    """G::Group::msm_unchecked(bases, scalars) truncated to the shorter side -> projective record."""
    if isinstance(bases, (list, tuple)):
        bases = b''.join(bytes(p) for p in bases)
    return bytes(lib.msm(np.frombuffer(bytes(bases), dtype=np.uint8), gp.fr_to_bytes(scalars), g2=group.g2))


def random_oracle_challenge(*chunks):
    """Stand-in for compute_random_oracle_challenge::<Fr, Blake2b512>: hash of the challenge contributions -> Fr."""
    h = hashlib.blake2b(digest_size=64)
    for c in chunks:
        h.update(bytes(c))
    return int.from_bytes(h.digest(), 'little') % R_MODULUS


def compressed(points, group=G1):
    """serialize_compressed of affine records (what challenge_contribution writes)."""
    return bytes(lib.serialize_points(np.frombuffer(bytes(points), dtype=np.uint8), g2=group.g2, compressed=True))


# ---- Schnorr (a12) -------------------------------------------------------------------------------------------------
class SchnorrCommitment:
    def __init__(self, bases, blindings, group=G1):
        """t = sum bases[i] * blindings[i]  (msm_unchecked; extra bases or blindings are ignored)."""
        self.group = group
        self.blindings = [b % R_MODULUS for b in blindings]
        self.t = gp.into_affine(_msm_unchecked(bases, self.blindings, group), group)

    @classmethod
    def new(cls, bases, blindings, group=G1):
        return cls(bases, blindings, group)

    def response(self, witnesses, challenge):
        if len(self.blindings) != len(witnesses):
            raise ValueError('ExpectedSameSizeSequences')
        return SchnorrResponse([(b + w * challenge) % R_MODULUS for b, w in zip(self.blindings, witnesses)], self.group)

    def challenge_contribution(self):
        return compressed(self.t, self.group)


class SchnorrResponse:
    def __init__(self, responses, group=G1):
        self.responses, self.group = list(responses), group

    def is_valid(self, bases, y, t, challenge):
        """bases[0]*responses[0] + ... + bases[n-1]*responses[n-1] - y*challenge == t"""
        nb = len(bases) if isinstance(bases, (list, tuple)) else len(bytes(bases)) // self.group.AFF
        if len(self.responses) != nb:
            raise ValueError('ExpectedSameSizeSequences')
        lhs = gp.add([_msm_unchecked(bases, self.responses, self.group), gp.mul(y, -challenge, self.group)], self.group)
        return gp.into_affine(lhs, self.group) == bytes(t)


class PokPedersenCommitmentProtocol:
    """Knowledge of (x1, x2) in base1*x1 + base2*x2 = Y (discrete_log.rs:178-220)."""

    def __init__(self, witness1, blinding1, base1, witness2, blinding2, base2):
        self.w1, self.b1, self.w2, self.b2 = witness1, blinding1, witness2, blinding2
        self.t = gp.into_affine(gp.add([gp.mul(base1, blinding1), gp.mul(base2, blinding2)]))

    init = classmethod(lambda cls, *a: cls(*a))

    def gen_proof(self, challenge):
        return PokPedersenCommitment(self.t, (self.b1 + self.w1 * challenge) % R_MODULUS, (self.b2 + self.w2 * challenge) % R_MODULUS)


@dataclass
class PokPedersenCommitment:
    t: bytes
    response1: int
    response2: int

    def verify(self, y, base1, base2, challenge):
        exp = gp.add([gp.mul(base1, self.response1), gp.mul(base2, self.response2), gp.mul(y, -challenge)])
        return gp.into_affine(exp) == self.t


# ---- BBS+ (a13 - a15) --------------------------------------------------------------------------------------------------
@dataclass
class SignatureParamsG1:
    g1: bytes
    g2: bytes
    h_0: bytes
    h: bytes                      # concatenated affine records h_1 .. h_n

    def supported_message_count(self):
        return len(self.h) // 96

    def h_at(self, i):
        return self.h[96 * i:96 * (i + 1)]

    def commit_to_messages(self, indexed_messages: Dict[int, int], blinding):
        """h_0 * blinding + sum h_i * m_i  (setup.rs:128-146): ONE msm_unchecked over the selected bases."""
        idx = sorted(indexed_messages)
        if idx and idx[-1] >= self.supported_message_count():
            raise ValueError('InvalidMessageIdx')
        bases = b''.join(self.h_at(i) for i in idx) + self.h_0
        scalars = [indexed_messages[i] for i in idx] + [blinding]
        return _msm_unchecked(bases, scalars)

    def b(self, indexed_messages: Dict[int, int], s):
        """g1 + h_0 * s + sum h_i * m_i  (setup.rs:150-163)."""
        return gp.add([self.commit_to_messages(indexed_messages, s), gp.to_projective(self.g1)])


@dataclass
class SignatureG1:
    A: bytes
    e: int
    s: int

    @classmethod
    def new(cls, messages: List[int], sk: int, params: SignatureParamsG1, e: int, s: int):
        """A = b * 1/(e + x)  (signature.rs:138-211)."""
        if len(messages) != params.supported_message_count():
            raise ValueError('MessageCountIncompatibleWithSigParams')
        b = gp.into_affine(params.b(dict(enumerate(messages)), s))
        inv = pow((e + sk) % R_MODULUS, -1, R_MODULUS)
        return cls(A=gp.mul_affine(b, inv), e=e % R_MODULUS, s=s % R_MODULUS)

    def verify(self, messages: List[int], pk: bytes, params: SignatureParamsG1):
        """e(A, pk + g2*e) == e(b, g2)  as the 2-pair product check (signature.rs:272-296)."""
        if len(messages) != params.supported_message_count():
            raise ValueError('MessageCountIncompatibleWithSigParams')
        if gp.is_identity(self.A):
            raise ValueError('ZeroSignature')
        b = gp.into_affine(params.b(dict(enumerate(messages)), self.s))
        g2_e_pk = gp.into_affine(gp.add([gp.mul(params.g2, self.e, G2), gp.to_projective(pk, G2)], G2), G2)
        return lib.multi_pairing_is_one(self.A + gp.neg(b), g2_e_pk + params.g2)


@dataclass
class PoKOfSignatureG1Proof:
    A_prime: bytes
    A_bar: bytes
    d: bytes
    sc_resp_1: PokPedersenCommitment
    T2: bytes
    sc_resp_2: SchnorrResponse

    def get_resp_for_message(self, msg_idx, revealed_msg_ids):
        """Schnorr response of hidden message msg_idx (proof.rs:448-476): its position among the hidden messages."""
        if msg_idx in revealed_msg_ids:
            raise ValueError('InvalidMsgIdxForResponse')
        pos = msg_idx - sum(1 for i in revealed_msg_ids if i < msg_idx)
        return self.sc_resp_2.responses[pos]

    def verify_schnorr_proofs(self, revealed_msgs: Dict[int, int], challenge, params: SignatureParamsG1):
        A_bar_minus_d = gp.into_affine(gp.add_affine([self.A_bar, gp.neg(self.d)]))
        if not self.sc_resp_1.verify(A_bar_minus_d, self.A_prime, params.h_0, challenge):
            return False                                 # FirstSchnorrVerificationFailed
        n = params.supported_message_count()
        bases_2 = b''.join(params.h_at(i) for i in range(n) if i not in revealed_msgs) + self.d + params.h_0
        rev = sorted(revealed_msgs)
        # pr = -g1 + sum_{i in D} h_i * (-m_i) = -(g1 + sum h_i m_i)
        parts = [gp.to_projective(params.g1)]
        if rev:
            parts.append(_msm_unchecked(b''.join(params.h_at(i) for i in rev), [revealed_msgs[i] for i in rev]))
        pr = gp.neg(gp.into_affine(gp.add(parts)))
        return self.sc_resp_2.is_valid(bases_2, pr, self.T2, challenge)

    def verify(self, revealed_msgs, challenge, pk, params: SignatureParamsG1):
        if gp.is_identity(self.A_prime):
            return False                                 # ZeroSignature
        if not self.verify_schnorr_proofs(revealed_msgs, challenge, params):
            return False
        return lib.multi_pairing_is_one(self.A_prime + gp.neg(self.A_bar), bytes(pk) + params.g2)

    def verify_with_randomized_pairing_checker(self, revealed_msgs, challenge, pk, params, pairing_checker):
        if gp.is_identity(self.A_prime) or not self.verify_schnorr_proofs(revealed_msgs, challenge, params):
            return False
        pairing_checker.add_sources(self.A_prime, bytes(pk), self.A_bar, params.g2)
        return True


class PoKOfSignatureG1Protocol:
    """proof.rs:159-251.  `revealed` = indices disclosed to the verifier; rnd = the prover's random scalars
    (r1, r2, blinding for -e, blinding for r2, one blinding per hidden message, blinding for -r3, blinding for s')."""

    def __init__(self, signature: SignatureG1, params: SignatureParamsG1, messages: List[int], revealed, rnd, blindings=None):
        """blindings: {message index: blinding} for MessageOrBlinding::BlindMessageWithConcreteBlinding (witness equalities
        across statements reuse one blinding); every other hidden message draws its blinding from rnd."""
        n = params.supported_message_count()
        if len(messages) != n:
            raise ValueError('MessageCountIncompatibleWithSigParams')
        rnd = iter(rnd)
        r1 = next(rnd) % R_MODULUS
        if r1 == 0:
            raise ValueError('r1 must be non-zero')
        r2 = next(rnd) % R_MODULUS
        r3 = pow(r1, -1, R_MODULUS)
        b = gp.into_affine(params.b(dict(enumerate(messages)), signature.s))
        A_prime = gp.mul_affine(signature.A, r1)
        b_r1 = gp.mul(b, r1)
        A_bar = gp.add([b_r1, gp.mul(A_prime, -signature.e)])
        d = gp.add([b_r1, gp.mul(params.h_0, -r2)])
        self.A_prime, self.A_bar, self.d = A_prime, gp.into_affine(A_bar), gp.into_affine(d)
        s_prime = (signature.s - r2 * r3) % R_MODULUS
        self.sc_comm_1 = PokPedersenCommitmentProtocol(-signature.e % R_MODULUS, next(rnd), A_prime, r2, next(rnd), params.h_0)
        hidden = [i for i in range(n) if i not in revealed]
        bases_2 = b''.join(params.h_at(i) for i in hidden) + self.d + params.h_0
        blindings = blindings or {}
        randomness_2 = [blindings[i] if i in blindings else next(rnd) for i in hidden] + [next(rnd), next(rnd)]
        self.sc_wits_2 = [messages[i] for i in hidden] + [-r3 % R_MODULUS, s_prime]
        self.sc_comm_2 = SchnorrCommitment.new(bases_2, randomness_2)

    init = classmethod(lambda cls, *a: cls(*a))

    def challenge_contribution(self, revealed_msgs: Dict[int, int], params: SignatureParamsG1):
        out = compressed(self.A_prime + self.A_bar + self.d + self.sc_comm_1.t + self.sc_comm_2.t)
        for i in sorted(revealed_msgs):
            out += i.to_bytes(8, 'little') + (revealed_msgs[i] % R_MODULUS).to_bytes(32, 'little')
        return out

    def gen_proof(self, challenge):
        return PoKOfSignatureG1Proof(A_prime=self.A_prime, A_bar=self.A_bar, d=self.d, sc_resp_1=self.sc_comm_1.gen_proof(challenge),
                                     T2=self.sc_comm_2.t, sc_resp_2=self.sc_comm_2.response(self.sc_wits_2, challenge))


# ---- VB accumulator (a16, a17) -----------------------------------------------------------------------------------------
def _poly_mul_linear(p, root):
    """p(x) * (root - x)"""
    out = [0] * (len(p) + 1)
    for i, c in enumerate(p):
        out[i] = (out[i] + c * root) % R_MODULUS
        out[i + 1] = (out[i + 1] - c) % R_MODULUS
    return out


def _poly_add_scaled(acc, p, k):
    if len(acc) < len(p):
        acc = acc + [0] * (len(p) - len(acc))
    for i, c in enumerate(p):
        acc[i] = (acc[i] + c * k) % R_MODULUS
    return acc


class Poly_d:
    @staticmethod
    def eval_direct(updates, x):
        acc = 1
        for y in updates:
            acc = (y - x) * acc % R_MODULUS
        return acc


class Poly_v_A:
    @staticmethod
    def generate(additions, alpha):
        """sum_s prod_{i<s}(y_i + alpha) * prod_{j>s}(y_j - x)  (batch_utils.rs:115-140) -> coefficient list."""
        n = len(additions)
        if n == 0:
            return []
        polys = [None] * n
        polys[n - 1] = [1]
        for s in range(1, n):
            polys[n - 1 - s] = _poly_mul_linear(polys[n - s], additions[n - s])
        acc, factor = [], 1
        for s in range(n):
            acc = _poly_add_scaled(acc, polys[s], factor)
            factor = factor * (additions[s] + alpha) % R_MODULUS
        return acc

    @staticmethod
    def factors(additions, alpha):
        """1, (y_0 + alpha), (y_0 + alpha)(y_1 + alpha), ...: independent of the evaluation point."""
        out, f = [], 1
        for a in additions:
            out.append(f)
            f = f * (a + alpha) % R_MODULUS
        return out

    @staticmethod
    def eval_direct(additions, alpha, x, factors=None):
        n = len(additions)
        if n == 0:
            return 0
        factors = factors or Poly_v_A.factors(additions, alpha)
        tot, poly = 0, 1
        for s in range(n - 1, -1, -1):                   # poly = prod_{j > s} (y_j - x)
            tot = (tot + factors[s] * poly) % R_MODULUS
            poly = poly * (additions[s] - x) % R_MODULUS
        return tot

    @staticmethod
    def eval_direct_on_batch(additions, alpha, xs):
        f = Poly_v_A.factors(additions, alpha)
        return [Poly_v_A.eval_direct(additions, alpha, x, f) for x in xs]


class Poly_v_D:
    @staticmethod
    def generate(removals, alpha):
        """sum_s prod_{i<=s} 1/(y_i + alpha) * prod_{j<s}(y_j - x)  (batch_utils.rs:270-296)."""
        n = len(removals)
        if n == 0:
            return []
        acc, poly, factor = [], [1], 1
        for s in range(n):
            factor = factor * pow((removals[s] + alpha) % R_MODULUS, -1, R_MODULUS) % R_MODULUS
            acc = _poly_add_scaled(acc, poly, factor)
            poly = _poly_mul_linear(poly, removals[s])
        return acc

    @staticmethod
    def factors(removals, alpha):
        """1/(y_0 + alpha), 1/((y_0 + alpha)(y_1 + alpha)), ...: independent of the evaluation point."""
        out, f = [], 1
        for d in removals:
            f = f * pow((d + alpha) % R_MODULUS, -1, R_MODULUS) % R_MODULUS
            out.append(f)
        return out

    @staticmethod
    def eval_direct(removals, alpha, x, factors=None):
        factors = factors or Poly_v_D.factors(removals, alpha)
        tot, poly = 0, 1
        for s in range(len(removals)):
            tot = (tot + factors[s] * poly) % R_MODULUS
            poly = poly * (removals[s] - x) % R_MODULUS
        return tot

    @staticmethod
    def eval_direct_on_batch(removals, alpha, xs):
        f = Poly_v_D.factors(removals, alpha)
        return [Poly_v_D.eval_direct(removals, alpha, x, f) for x in xs]


class Poly_v_AD:
    @staticmethod
    def compute_factor(additions, alpha):
        f = 1
        for a in additions:
            f = f * (a + alpha) % R_MODULUS
        return f

    @staticmethod
    def generate(additions, removals, alpha):
        p = Poly_v_A.generate(additions, alpha)
        if removals:
            p = _poly_add_scaled(p, Poly_v_D.generate(removals, alpha), -Poly_v_AD.compute_factor(additions, alpha) % R_MODULUS)
        return p

    @staticmethod
    def eval_direct(additions, removals, alpha, x):
        e = Poly_v_A.eval_direct(additions, alpha, x)
        if removals:
            e = (e - Poly_v_D.eval_direct(removals, alpha, x) * Poly_v_AD.compute_factor(additions, alpha)) % R_MODULUS
        return e

    @staticmethod
    def eval_direct_on_batch(additions, removals, alpha, xs):
        f = Poly_v_AD.compute_factor(additions, alpha)
        a = Poly_v_A.eval_direct_on_batch(additions, alpha, xs)
        if removals:
            b = Poly_v_D.eval_direct_on_batch(removals, alpha, xs)
            a = [(x - y * f) % R_MODULUS for x, y in zip(a, b)]
        return a


class Omega:
    """c_0 * V, c_1 * V, ... for the coefficients of v_AD (batch_utils.rs:498-509): fixed-base multiplication of V by
    every coefficient + normalize_batch, fused on the device."""

    def __init__(self, points: bytes):
        self.points = bytes(points)

    @classmethod
    def new(cls, additions, removals, old_accumulator, sk):
        coeffs = Poly_v_AD.generate(additions, removals, sk)
        tbl = lib.FixedBaseTable(old_accumulator, len(coeffs))             # multiply_field_elems_with_same_group_elem
        try:
            return cls(bytes(tbl.mul_many_normalized(gp.fr_to_bytes(coeffs))))
        finally:
            tbl.free()

    def __len__(self):
        return len(self.points) // 96

    @staticmethod
    def scaled_powers_of_y(y, scalar, n):
        out = []
        cur = scalar % R_MODULUS
        for _ in range(n):
            out.append(cur)
            cur = cur * y % R_MODULUS
        return out

    def evaluate(self, y, scalar):
        """<scalar * powers of y, omega>: ONE msm_unchecked (batch_utils.rs:664-668)."""
        return _msm_unchecked(self.points, self.scaled_powers_of_y(y, scalar, len(self)))


def compute_update_using_secret_key_after_batch_updates(additions, removals, elements, old_witnesses, old_accumulator, sk):
    """witness.rs:243-284: new C_i = d_A(y_i)/d_D(y_i) * C_i + v_AD(y_i)/d_D(y_i) * V for every element, as ONE fused
    device call (independent scalar multiplications + fixed-base multiplications of V + normalize_batch).
    -> (d_factors, new witnesses as concatenated affine records)"""
    m = len(elements)
    if m * 96 != len(bytes(old_witnesses)):
        raise ValueError('NeedSameNoOfElementsAndWitnesses')
    d_factors, v_factors = [], []
    v_AD = Poly_v_AD.eval_direct_on_batch(additions, removals, sk, elements)
    for y, v in zip(elements, v_AD):
        d_D_inv = pow(Poly_d.eval_direct(removals, y), -1, R_MODULUS)
        d_factors.append(Poly_d.eval_direct(additions, y) * d_D_inv % R_MODULUS)
        v_factors.append(v * d_D_inv % R_MODULUS)
    wits = np.frombuffer(bytes(old_witnesses), dtype=np.uint8)
    if m < 65536:
        # at the reference's batch sizes the device shares one doubling chain between d * C_i and v * V instead of building
        # WindowTable::new(m, V) first (dg_batch_mul_add_same_g1)
        out = lib.batch_mul_add_same_g1(wits, gp.fr_to_bytes(d_factors), old_accumulator, gp.fr_to_bytes(v_factors))
    else:
        table = msm_mod.WindowTable.new(m, old_accumulator)
        try:
            out = lib.batch_mul_add_fixed_g1(wits, gp.fr_to_bytes(d_factors), table._t, gp.fr_to_bytes(v_factors))
        finally:
            table.free()
    return d_factors, bytes(out)


def compute_update_using_public_info_after_batch_updates(additions, removals, omega: Omega, element, old_witness):
    """witness.rs:290-345: the holder's update without the secret key: C' = d_A/d_D * C + 1/d_D * <powers of y, omega>."""
    d_A = Poly_d.eval_direct(additions, element)
    d_D_inv = pow(Poly_d.eval_direct(removals, element), -1, R_MODULUS)
    y_omega_ip = omega.evaluate(element, d_D_inv)
    return gp.into_affine(gp.add([gp.mul(old_witness, d_A * d_D_inv), y_omega_ip]))


def verify_membership_given_accumulated(V, member, witness, pk_Q_tilde, P_tilde):
    """e(witness, member * P_tilde + Q_tilde) * e(V, -P_tilde) == 1  (positive.rs:401-425)."""
    rhs = gp.into_affine(gp.add([gp.mul(P_tilde, member, G2), gp.to_projective(pk_Q_tilde, G2)], G2), G2)
    return lib.multi_pairing_is_one(bytes(witness) + bytes(V), rhs + gp.neg(P_tilde, G2))
