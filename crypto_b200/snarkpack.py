"""Host-side mirror of SnarkPack aggregation of Groth16 proofs over the C ABI.

Same structure, names and relations as legogroth16/src/aggregation/:
  * srs.rs:272-428          setup_fake_srs, GenericSRS::specialize -> (ProverSRS, VerifierSRS)
  * key.rs:12-187           Key (VKey over G2, WKey over G1): split / scale / compress / first
  * commitment.rs:15-70     PairCommitment::{single, double}
  * utils.rs:26-160         compress, inner_product_and_{single,double}_commitments, aggregate_public_inputs,
                            prove_commitments, verify_kzg, final_verification_check
  * kzg.rs:32-343           KZG openings of the final commitment keys (product-form polynomials of the GIPA challenges)
  * groth16/prover.rs:46-382    aggregate_proofs, prove_tipp_mipp, gipa_tipp_mipp
  * groth16/verifier.rs:34-454  verify_aggregate_proof, verify_tipp_mipp, gipa_verify_tipp_mipp
  * legogroth16/prover.rs:47-424, verifier.rs:48-330   aggregate_lego_proofs / verify_aggregate_lego_proof: the same
                            protocol with a second MIPP for the D commitments (com_d, z_d, comms_d, final_d)
  * legogroth16/using_groth16.rs:27-128   LegoGroth16 proofs through the Groth16 aggregate with the D's in the clear
The Fiat-Shamir transcript is a hash chain (the reference's Merlin transcript is outside the hot path); everything that
costs curve arithmetic runs on the GPU: per GIPA round TEN pairing products go out as ONE dg_multi_pairing_batch call,
the vector foldings are dg_compress_g1/g2, key scaling is dg_batch_mul, the MIPP inner products and the KZG quotient
commitments are MSMs, and the verifier's pairing equations meet in one RandomizedPairingChecker (one final exponentiation).

Group elements are packed Montgomery records (bytes); vectors are lists of records.
"""
import hashlib
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import group as gp
from . import lib
from . import pairing_check as pc
from .group import G1, G2, R_MODULUS


# ---- transcript -----------------------------------------------------------------------------------------------------
class Transcript:
    def __init__(self, label: bytes):
        self.state = hashlib.blake2b(b'snarkpack' + label, digest_size=64).digest()

    def append(self, label: bytes, data):
        if isinstance(data, int):
            data = (data % R_MODULUS).to_bytes(32, 'little')
        if isinstance(data, (tuple, list)):
            data = b''.join(bytes(d) for d in data)
        self.state = hashlib.blake2b(self.state + label + bytes(data), digest_size=64).digest()

    def challenge_scalar(self, label: bytes) -> int:
        ctr = 0
        while True:
            h = hashlib.blake2b(self.state + label + ctr.to_bytes(4, 'little'), digest_size=64).digest()
            c = int.from_bytes(h, 'little') % R_MODULUS
            if c:
                self.state = h
                return c
            ctr += 1


def powers(x, n):
    out, cur = [], 1
    for _ in range(n):
        out.append(cur)
        cur = cur * x % R_MODULUS
    return out


def _inv(x):
    return pow(x % R_MODULUS, -1, R_MODULUS)


def _cat(v):
    return np.frombuffer(b''.join(bytes(x) for x in v), dtype=np.uint8)


def _split_records(buf, size):
    b = bytes(buf)
    return [b[i:i + size] for i in range(0, len(b), size)]


# ---- SRS --------------------------------------------------------------------------------------------------------------
@dataclass
class GenericSRS:
    g_alpha_powers: List[bytes]
    h_alpha_powers: List[bytes]
    g_beta_powers: List[bytes]
    h_beta_powers: List[bytes]

    def specialize(self, num_proofs):
        n = num_proofs
        assert n & (n - 1) == 0 and len(self.g_alpha_powers) >= 2 * n
        vkey = Key(self.h_alpha_powers[:n], self.h_beta_powers[:n], G2)
        wkey = Key(self.g_alpha_powers[n:2 * n], self.g_beta_powers[n:2 * n], G1)
        pk = ProverSRS(n, self.g_alpha_powers[:2 * n], self.h_alpha_powers[:n], self.g_beta_powers[:2 * n], self.h_beta_powers[:n], vkey, wkey)
        vk = VerifierSRS(n, self.g_alpha_powers[0], self.h_alpha_powers[0], self.g_alpha_powers[1], self.g_beta_powers[1],
                         self.h_alpha_powers[1], self.h_beta_powers[1])
        return pk, vk


def structured_generators_scalar_power(num, g, s, group=G1):
    """g, g^s, g^(s^2), ...: fixed-base multiplication of g by the powers of s + normalize_batch, fused on the device."""
    tbl = lib.FixedBaseTable(g, num, g2=group.g2)
    try:
        return _split_records(tbl.mul_many_normalized(gp.fr_to_bytes(powers(s, num))), group.AFF)
    finally:
        tbl.free()


def setup_fake_srs(alpha, beta, size, g, h):
    return GenericSRS(structured_generators_scalar_power(2 * size, g, alpha), structured_generators_scalar_power(2 * size, h, alpha, G2),
                      structured_generators_scalar_power(2 * size, g, beta), structured_generators_scalar_power(2 * size, h, beta, G2))


@dataclass
class Key:
    a: List[bytes]
    b: List[bytes]
    group: type

    def __len__(self):
        return len(self.a)

    def split(self, at):
        return Key(self.a[:at], self.b[:at], self.group), Key(self.a[at:], self.b[at:], self.group)

    def scale(self, s_vec):
        if len(self.a) != len(s_vec):
            raise ValueError('InvalidKeyLength')
        sc = gp.fr_to_bytes(s_vec)
        g = self.group
        a = lib.normalize_batch(lib.batch_mul(_cat(self.a), sc, g2=g.g2), g2=g.g2)
        b = lib.normalize_batch(lib.batch_mul(_cat(self.b), sc, g2=g.g2), g2=g.g2)
        return Key(_split_records(a, g.AFF), _split_records(b, g.AFF), g)

    def compress(self, right, scale):
        if len(self.a) != len(right.a):
            raise ValueError('InvalidKeyLength')
        g, sc = self.group, gp.fr_to_bytes([scale])
        return Key(_split_records(lib.compress(_cat(self.a), _cat(right.a), sc, g2=g.g2), g.AFF),
                   _split_records(lib.compress(_cat(self.b), _cat(right.b), sc, g2=g.g2), g.AFF), g)

    def first(self):
        return self.a[0], self.b[0]


@dataclass
class ProverSRS:
    n: int
    g_alpha_powers_table: List[bytes]
    h_alpha_powers_table: List[bytes]
    g_beta_powers_table: List[bytes]
    h_beta_powers_table: List[bytes]
    vkey: Key
    wkey: Key

    def has_correct_len(self, n):
        return len(self.vkey) == n and len(self.wkey) == n


@dataclass
class VerifierSRS:
    n: int
    g: bytes
    h: bytes
    g_alpha: bytes
    g_beta: bytes
    h_alpha: bytes
    h_beta: bytes


# ---- pairing commitments -------------------------------------------------------------------------------------------
def _pairing_products(products):
    """[(g1 records, g2 records), ...] -> one GT record per product, ONE device call for all of them."""
    g1 = _cat([p for a, _ in products for p in a])
    g2 = _cat([q for _, b in products for q in b])
    return [bytes(x) for x in lib.multi_pairing_batch(g1, g2, [len(a) for a, _ in products])]


def commit_single_products(vkey: Key, a_vec):
    """PairCommitment::single as two pairing products: T = prod e(A_i, v1_i), U = prod e(A_i, v2_i)."""
    if len(vkey) < len(a_vec):
        raise ValueError('InsufficientKeyLength')
    m = len(a_vec)
    return [(a_vec, vkey.a[:m]), (a_vec, vkey.b[:m])]


def commit_double_products(vkey: Key, wkey: Key, a, b):
    """PairCommitment::double: T = prod e(A_i, v1_i) prod e(w1_i, B_i), U likewise with (v2, w2)."""
    m, k = len(a), len(b)
    return [(list(a) + wkey.a[:k], vkey.a[:m] + list(b)), (list(a) + wkey.b[:k], vkey.b[:m] + list(b))]


def compress(vec, split, scalar, group=G1):
    """utils::compress: vec[i] + vec[i + split] * scalar for i < split."""
    out = lib.compress(_cat(vec[:split]), _cat(vec[split:2 * split]), gp.fr_to_bytes([scalar]), g2=group.g2)
    return _split_records(out, group.AFF)


# ---- KZG openings of the final keys ------------------------------------------------------------------------------------
def polynomial_evaluation_product_form_from_transcript(transcript, z, r_shift):
    power_zr = z * r_shift % R_MODULUS
    res = (1 + transcript[0] * power_zr) % R_MODULUS
    for x in transcript[1:]:
        power_zr = power_zr * power_zr % R_MODULUS
        res = res * (1 + x * power_zr) % R_MODULUS
    return res


def polynomial_coefficients_from_transcript(transcript, r_shift):
    coefficients, power_2_r = [1], r_shift % R_MODULUS
    for i, x in enumerate(transcript):
        if i > 0:
            power_2_r = power_2_r * power_2_r % R_MODULUS
        k = x * power_2_r % R_MODULUS
        coefficients += [c * k % R_MODULUS for c in coefficients]
    return coefficients


def create_kzg_opening(table_alpha, table_beta, coeffs, eval_poly, z, group):
    """(poly - eval) / (X - z) committed under both tables: two MSMs."""
    if len(coeffs) != len(table_alpha):
        raise ValueError('InvalidSRS: SRS len %d != coefficients len %d' % (len(table_alpha), len(coeffs)))
    num = list(coeffs)
    num[0] = (num[0] - eval_poly) % R_MODULUS
    q, carry = [0] * len(num), 0
    for i in range(len(num) - 1, 0, -1):                    # synthetic division by (X - z); the remainder is dropped
        carry = (num[i] + carry * z) % R_MODULUS
        q[i - 1] = carry
    sc = gp.fr_to_bytes(q)
    a = gp.into_affine(bytes(lib.msm(_cat(table_alpha), sc, g2=group.g2)), group)
    b = gp.into_affine(bytes(lib.msm(_cat(table_beta), sc, g2=group.g2)), group)
    return a, b


def prove_commitment_v(table_alpha, table_beta, transcript, z):
    coeffs = polynomial_coefficients_from_transcript(transcript, 1)
    return create_kzg_opening(table_alpha, table_beta, coeffs, polynomial_evaluation_product_form_from_transcript(transcript, z, 1), z, G2)


def prove_commitment_w(table_alpha, table_beta, transcript, r_shift, z):
    f = polynomial_coefficients_from_transcript(transcript, r_shift)
    n = len(f)
    fw = [0] * n + f                                           # X^n f(X)
    fwz = polynomial_evaluation_product_form_from_transcript(transcript, z, r_shift) * pow(z, n, R_MODULUS) % R_MODULUS
    return create_kzg_opening(table_alpha, table_beta, fw, fwz, z, G1)


# ---- prover ---------------------------------------------------------------------------------------------------------------
@dataclass
class GipaProof:
    nproofs: int
    comms_ab: list
    comms_c: list
    z_ab: list
    z_c: list
    final_a: bytes
    final_b: bytes
    final_c: bytes
    final_vkey: Tuple[bytes, bytes]
    final_wkey: Tuple[bytes, bytes]
    # LegoGroth16 aggregation (GipaProofLego, aggregation/legogroth16/proof.rs:79-95): the same MIPP once more for the D's
    comms_d: Optional[list] = None
    z_d: Optional[list] = None
    final_d: Optional[bytes] = None


@dataclass
class AggregateProof:
    """AggregateProof (aggregation/groth16/proof.rs) or, with com_d / z_d set, AggregateLegoProof
    (aggregation/legogroth16/proof.rs:22-42)."""
    com_ab: Tuple[bytes, bytes]
    com_c: Tuple[bytes, bytes]
    z_ab: bytes
    z_c: bytes
    gipa: GipaProof
    vkey_opening: Tuple[bytes, bytes]
    wkey_opening: Tuple[bytes, bytes]
    com_d: Optional[Tuple[bytes, bytes]] = None
    z_d: Optional[bytes] = None

    @property
    def is_lego(self):
        return self.com_d is not None


def gipa_tipp_mipp(transcript, a, b, c, vkey: Key, wkey: Key, r, ip_ab, agg_c, d=None, agg_d=None):
    """GIPA recursion for TIPP (A, B) and MIPP (C, r) -- aggregation/groth16/prover.rs:196-357 -- and, when d is given, the
    second MIPP (D, r) of the LegoGroth16 variant (aggregation/legogroth16/prover.rs:216-424): per round the ten (fourteen)
    pairing products go to the GPU in one call, the z values are two (four) MSMs."""
    lego = d is not None
    m_a, m_b, m_c, m_r = list(a), list(b), list(c), list(r)
    m_d = list(d) if lego else None
    comms_ab, comms_c, z_ab, z_c, challenges, challenges_inv = [], [], [], [], [], []
    comms_d, z_d = [], []
    transcript.append(b'inner-product-ab', ip_ab)
    transcript.append(b'comm-c', agg_c)
    if lego:
        transcript.append(b'comm-d', agg_d)
    c_inv = transcript.challenge_scalar(b'first-challenge')
    ch = _inv(c_inv)
    i = 0
    while len(m_a) > 1:
        split = len(m_a) // 2
        a_l, a_r, b_l, b_r = m_a[:split], m_a[split:], m_b[:split], m_b[split:]
        c_l, c_r, r_l, r_r = m_c[:split], m_c[split:], m_r[:split], m_r[split:]
        vk_l, vk_r = vkey.split(split)
        wk_l, wk_r = wkey.split(split)
        # TIPP: tab_l = double(vk_l, wk_r, a_r, b_l), tab_r = double(vk_r, wk_l, a_l, b_r), zab_l = e(a_r, b_l), zab_r = e(a_l, b_r)
        # MIPP: tuc_l = single(vk_l, c_r), tuc_r = single(vk_r, c_l): ten pairing products, one device call
        prods = (commit_double_products(vk_l, wk_r, a_r, b_l) + commit_double_products(vk_r, wk_l, a_l, b_r) +
                 [(a_r, b_l), (a_l, b_r)] + commit_single_products(vk_l, c_r) + commit_single_products(vk_r, c_l))
        if lego:
            d_l, d_r = m_d[:split], m_d[split:]
            prods += commit_single_products(vk_l, d_r) + commit_single_products(vk_r, d_l)
        gt = _pairing_products(prods)
        tab_l, tab_r, zab_l, zab_r, tuc_l, tuc_r = (gt[0], gt[1]), (gt[2], gt[3]), gt[4], gt[5], (gt[6], gt[7]), (gt[8], gt[9])
        zc_l = gp.into_affine(bytes(lib.msm(_cat(c_r), gp.fr_to_bytes(r_l))))      # c[n':] ^ r[:n']
        zc_r = gp.into_affine(bytes(lib.msm(_cat(c_l), gp.fr_to_bytes(r_r))))      # c[:n'] ^ r[n':]
        if lego:
            tud_l, tud_r = (gt[10], gt[11]), (gt[12], gt[13])
            zd_l = gp.into_affine(bytes(lib.msm(_cat(d_r), gp.fr_to_bytes(r_l))))
            zd_r = gp.into_affine(bytes(lib.msm(_cat(d_l), gp.fr_to_bytes(r_r))))
        if i > 0:
            transcript.append(b'c_inv', c_inv)
            labelled = [(b'zab_l', zab_l), (b'zab_r', zab_r), (b'zc_l', zc_l), (b'zc_r', zc_r)]
            if lego:
                labelled += [(b'zd_l', zd_l), (b'zd_r', zd_r)]
            labelled += [(b'tab_l', tab_l), (b'tab_r', tab_r), (b'tuc_l', tuc_l), (b'tuc_r', tuc_r)]
            if lego:
                labelled += [(b'tud_l', tud_l), (b'tud_r', tud_r)]
            for label, val in labelled:
                transcript.append(label, val)
            c_inv = transcript.challenge_scalar(b'challenge_i')
            ch = _inv(c_inv)
        m_a = compress(m_a, split, ch)
        m_b = compress(m_b, split, c_inv, G2)
        m_c = compress(m_c, split, ch)
        if lego:
            m_d = compress(m_d, split, ch)
            comms_d.append((tud_l, tud_r)); z_d.append((zd_l, zd_r))
        m_r = [(x + y * c_inv) % R_MODULUS for x, y in zip(r_l, r_r)]
        vkey = vk_l.compress(vk_r, c_inv)
        wkey = wk_l.compress(wk_r, ch)
        comms_ab.append((tab_l, tab_r)); comms_c.append((tuc_l, tuc_r))
        z_ab.append((zab_l, zab_r)); z_c.append((zc_l, zc_r))
        challenges.append(ch); challenges_inv.append(c_inv)
        i += 1
    proof = GipaProof(len(a), comms_ab, comms_c, z_ab, z_c, m_a[0], m_b[0], m_c[0], vkey.first(), wkey.first())
    if lego:
        proof.comms_d, proof.z_d, proof.final_d = comms_d, z_d, m_d[0]
    return proof, challenges, challenges_inv


def aggregate_proofs(srs: ProverSRS, transcript: Transcript, proofs):
    """proofs: list of (A, B, C) Groth16 proofs (affine records); the count must be a power of two >= 2
    (aggregation/groth16/prover.rs:46-148)."""
    return _aggregate(srs, transcript, proofs, False)


def aggregate_lego_proofs(srs: ProverSRS, transcript: Transcript, proofs):
    """proofs: list of (A, B, C, D) LegoGroth16 proofs; the D's get their own pair commitment, z_d = sum r^i D_i and MIPP
    (aggregation/legogroth16/prover.rs:47-156)."""
    return _aggregate(srs, transcript, proofs, True)


def aggregate_lego_proofs_using_groth16(srs: ProverSRS, transcript: Transcript, proofs):
    """LegoGroth16 proofs aggregated with the Groth16 protocol, the D's handed to the verifier in the clear
    (aggregation/legogroth16/using_groth16.rs:27-45) -> (AggregateProof, [D_i])."""
    return _aggregate(srs, transcript, [p[:3] for p in proofs], False), [p[3] for p in proofs]


def _aggregate(srs: ProverSRS, transcript: Transcript, proofs, lego):
    n = len(proofs)
    if n < 2:
        raise ValueError('InvalidProof: invalid proof size < 2')
    if n & (n - 1):
        raise ValueError('InvalidProof: invalid proof size: not power of two')
    if not srs.has_correct_len(n):
        raise ValueError('InvalidSRS: SRS len %d != proofs len %d' % (len(srs.vkey), n))
    a, b, c = [p[0] for p in proofs], [p[1] for p in proofs], [p[2] for p in proofs]
    d = [p[3] for p in proofs] if lego else None
    gt = _pairing_products(commit_double_products(srs.vkey, srs.wkey, a, b) + commit_single_products(srs.vkey, c) +
                           (commit_single_products(srs.vkey, d) if lego else []))
    com_ab, com_c = (gt[0], gt[1]), (gt[2], gt[3])
    com_d = (gt[4], gt[5]) if lego else None
    transcript.append(b'AB-commitment', com_ab)
    transcript.append(b'C-commitment', com_c)
    if lego:
        transcript.append(b'D-commitment', com_d)
    r = transcript.challenge_scalar(b'r-random-fiatshamir')
    r_vec = powers(r, n)
    r_inv = [_inv(x) for x in r_vec]
    b_r = _split_records(lib.normalize_batch(lib.batch_mul(_cat(b), gp.fr_to_bytes(r_vec), g2=True), g2=True), 192)   # B^r
    z_ab = _pairing_products([(a, b_r)])[0]
    z_c = gp.into_affine(bytes(lib.msm(_cat(c), gp.fr_to_bytes(r_vec))))
    z_d = gp.into_affine(bytes(lib.msm(_cat(d), gp.fr_to_bytes(r_vec)))) if lego else None
    wkey_r_inv = srs.wkey.scale(r_inv)
    gipa, challenges, challenges_inv = gipa_tipp_mipp(transcript, a, b_r, c, srs.vkey, wkey_r_inv, r_vec, z_ab, z_c, d, z_d)
    challenges.reverse()
    challenges_inv.reverse()
    r_inverse = _inv(r_vec[1])
    transcript.append(b'kzg-challenge', challenges[0])
    transcript.append(b'vkey0', gipa.final_vkey[0]); transcript.append(b'vkey1', gipa.final_vkey[1])
    transcript.append(b'wkey0', gipa.final_wkey[0]); transcript.append(b'wkey1', gipa.final_wkey[1])
    z = transcript.challenge_scalar(b'z-challenge')
    vkey_opening = prove_commitment_v(srs.h_alpha_powers_table, srs.h_beta_powers_table, challenges_inv, z)
    wkey_opening = prove_commitment_w(srs.g_alpha_powers_table, srs.g_beta_powers_table, challenges, r_inverse, z)
    return AggregateProof(com_ab, com_c, z_ab, z_c, gipa, vkey_opening, wkey_opening, com_d, z_d)


# ---- verifier ---------------------------------------------------------------------------------------------------------------
def gipa_verify_tipp_mipp(proof: AggregateProof, r_shift, transcript):
    gipa = proof.gipa
    lego = proof.is_lego
    challenges, challenges_inv = [], []
    transcript.append(b'inner-product-ab', proof.z_ab)
    transcript.append(b'comm-c', proof.z_c)
    if lego:
        transcript.append(b'comm-d', proof.z_d)
    c_inv = transcript.challenge_scalar(b'first-challenge')
    ch = _inv(c_inv)
    for i, ((tab_l, tab_r), (zab_l, zab_r), (tuc_l, tuc_r), (zc_l, zc_r)) in enumerate(zip(gipa.comms_ab, gipa.z_ab, gipa.comms_c, gipa.z_c)):
        if i > 0:
            transcript.append(b'c_inv', c_inv)
            labelled = [(b'zab_l', zab_l), (b'zab_r', zab_r), (b'zc_l', zc_l), (b'zc_r', zc_r)]
            if lego:
                labelled += [(b'zd_l', gipa.z_d[i][0]), (b'zd_r', gipa.z_d[i][1])]
            labelled += [(b'tab_l', tab_l), (b'tab_r', tab_r), (b'tuc_l', tuc_l), (b'tuc_r', tuc_r)]
            if lego:
                labelled += [(b'tud_l', gipa.comms_d[i][0]), (b'tud_r', gipa.comms_d[i][1])]
            for label, val in labelled:
                transcript.append(label, val)
            c_inv = transcript.challenge_scalar(b'challenge_i')
            ch = _inv(c_inv)
        challenges.append(ch)
        challenges_inv.append(c_inv)
    res = {'tab': proof.com_ab[0], 'uab': proof.com_ab[1], 'zab': proof.z_ab, 'tc': proof.com_c[0], 'uc': proof.com_c[1]}
    # zc += <(zc_l, zc_r) per round, (c, c_inv) per round>: one MSM
    zc_b = [p for pair in gipa.z_c for p in pair]
    z_s = [s for pair in zip(challenges, challenges_inv) for s in pair]
    zc = gp.add([gp.to_projective(proof.z_c), bytes(lib.msm(_cat(zc_b), gp.fr_to_bytes(z_s)))])
    for (tab_l, tab_r), (zab_l, zab_r), (tuc_l, tuc_r), c_, ci_ in zip(gipa.comms_ab, gipa.z_ab, gipa.comms_c, challenges, challenges_inv):
        for key, left, right in (('tab', tab_l[0], tab_r[0]), ('uab', tab_l[1], tab_r[1]), ('zab', zab_l, zab_r),
                                 ('tc', tuc_l[0], tuc_r[0]), ('uc', tuc_l[1], tuc_r[1])):
            res[key] = pc.gt_add(res[key], pc.gt_add(pc.gt_mul_bigint(left, c_), pc.gt_mul_bigint(right, ci_)))
    res['zc'] = gp.into_affine(zc)
    if lego:
        res['td'], res['ud'] = proof.com_d
        zd_b = [p for pair in gipa.z_d for p in pair]
        res['zd'] = gp.into_affine(gp.add([gp.to_projective(proof.z_d), bytes(lib.msm(_cat(zd_b), gp.fr_to_bytes(z_s)))]))
        for (tud_l, tud_r), c_, ci_ in zip(gipa.comms_d, challenges, challenges_inv):
            for key, left, right in (('td', tud_l[0], tud_r[0]), ('ud', tud_l[1], tud_r[1])):
                res[key] = pc.gt_add(res[key], pc.gt_add(pc.gt_mul_bigint(left, c_), pc.gt_mul_bigint(right, ci_)))
    challenges.reverse()
    challenges_inv.reverse()
    final_r = polynomial_evaluation_product_form_from_transcript(challenges_inv, r_shift, 1)
    return res, final_r, challenges, challenges_inv


def verify_kzg(v: VerifierSRS, final_vkey, vkey_opening, final_wkey, wkey_opening, challenges, challenges_inv, shift, z, checker):
    zero = pc.gt_zero()
    vpoly = polynomial_evaluation_product_form_from_transcript(challenges_inv, z, 1)
    ng = gp.neg(v.g)
    for cf, vk, pi in ((final_vkey[0], v.g_alpha, vkey_opening[0]), (final_vkey[1], v.g_beta, vkey_opening[1])):
        b = gp.into_affine(gp.add([gp.to_projective(cf, G2), gp.mul(v.h, -vpoly, G2)], G2), G2)          # cf - h^y
        c = gp.into_affine(gp.add([gp.to_projective(vk), gp.mul(v.g, -z)]))                              # g^alpha - g^z
        checker.add_multiple_sources_and_target([ng, c], [b, pi], zero)
    fwz = polynomial_evaluation_product_form_from_transcript(challenges, z, shift) * pow(z, v.n, R_MODULUS) % R_MODULUS
    nh = gp.neg(v.h, G2)
    for cf, wk, pi in ((final_wkey[0], v.h_alpha, wkey_opening[0]), (final_wkey[1], v.h_beta, wkey_opening[1])):
        a = gp.into_affine(gp.add([gp.to_projective(cf), gp.mul(v.g, -fwz)]))                            # cf - g^y
        d = gp.into_affine(gp.add([gp.to_projective(wk, G2), gp.mul(v.h, -z, G2)], G2), G2)              # h^alpha - h^z
        checker.add_multiple_sources_and_target([a, pi], [nh, d], zero)


def verify_tipp_mipp(v: VerifierSRS, proof: AggregateProof, r_shift, transcript, checker):
    final_res, final_r, challenges, challenges_inv = gipa_verify_tipp_mipp(proof, r_shift, transcript)
    g = proof.gipa
    transcript.append(b'kzg-challenge', challenges[0])
    transcript.append(b'vkey0', g.final_vkey[0]); transcript.append(b'vkey1', g.final_vkey[1])
    transcript.append(b'wkey0', g.final_wkey[0]); transcript.append(b'wkey1', g.final_wkey[1])
    z = transcript.challenge_scalar(b'z-challenge')
    verify_kzg(v, g.final_vkey, proof.vkey_opening, g.final_wkey, proof.wkey_opening, challenges, challenges_inv, _inv(r_shift), z, checker)
    checker.add_multiple_sources_and_target([g.final_a], [g.final_b], final_res['zab'])
    checker.add_multiple_sources_and_target([g.final_a, g.final_wkey[0]], [g.final_vkey[0], g.final_b], final_res['tab'])
    checker.add_multiple_sources_and_target([g.final_a, g.final_wkey[1]], [g.final_vkey[1], g.final_b], final_res['uab'])
    checker.add_multiple_sources_and_target([g.final_c], [g.final_vkey[0]], final_res['tc'])
    checker.add_multiple_sources_and_target([g.final_c], [g.final_vkey[1]], final_res['uc'])
    ok = gp.mul_affine(g.final_c, final_r) == final_res['zc']
    if proof.is_lego:                                           # MIPP for D (aggregation/legogroth16/verifier.rs:194-223)
        checker.add_multiple_sources_and_target([g.final_d], [g.final_vkey[0]], final_res['td'])
        checker.add_multiple_sources_and_target([g.final_d], [g.final_vkey[1]], final_res['ud'])
        ok = ok and gp.mul_affine(g.final_d, final_r) == final_res['zd']
    return ok


def aggregate_public_inputs(public_inputs, r_powers, r_sum, gamma_abc_g1: bytes):
    """S_0^(sum r^j) * prod_i S_i^(sum_j a_(j,i) r^j): one MSM over gamma_abc_g1 (utils.rs:112-140)."""
    l = len(public_inputs[0])
    scalars = [r_sum] + [sum(inp[i] * rj for inp, rj in zip(public_inputs, r_powers)) % R_MODULUS for i in range(l)]
    return gp.into_affine(bytes(lib.msm(np.frombuffer(gamma_abc_g1, dtype=np.uint8), gp.fr_to_bytes(scalars))))


def verify_aggregate_lego_proof(v: VerifierSRS, vk, public_inputs, proof: AggregateProof, transcript: Transcript, checker_random, lazy=True):
    """aggregation/legogroth16/verifier.rs:48-112: the D commitment joins the transcript, the D MIPP joins the checker and
    z_d pairs with gamma in the final product."""
    if not proof.is_lego or proof.gipa.comms_d is None or proof.gipa.z_d is None or proof.gipa.final_d is None:
        return False
    return verify_aggregate_proof(v, vk, public_inputs, proof, transcript, checker_random, lazy)


def verify_aggregate_lego_proof_using_groth16(v: VerifierSRS, vk, public_inputs, proof: AggregateProof, d, transcript: Transcript,
                                              checker_random, lazy=True):
    """aggregation/legogroth16/using_groth16.rs:47-128: a Groth16 aggregate plus the D's in the clear; the verifier folds
    sum r^i D_i into the public-input term itself."""
    if proof.is_lego or len(d) != proof.gipa.nproofs:
        return False
    return verify_aggregate_proof(v, vk, public_inputs, proof, transcript, checker_random, lazy, clear_d=d)


def verify_aggregate_proof(v: VerifierSRS, vk, public_inputs, proof: AggregateProof, transcript: Transcript, checker_random, lazy=True,
                           clear_d=None):
    """vk: crypto_b200.groth16.VerifyingKey of the aggregated circuit; public_inputs: one list per proof."""
    n = proof.gipa.nproofs
    rounds = n.bit_length() - 1
    vectors = [proof.gipa.comms_ab, proof.gipa.comms_c, proof.gipa.z_ab, proof.gipa.z_c]
    if proof.is_lego:
        vectors += [proof.gipa.comms_d, proof.gipa.z_d]
    if n < 2 or n & (n - 1) or any(x is None or len(x) != rounds for x in vectors):
        return False                                            # parsing_check
    if any(len(p) + 1 > len(vk.gamma_abc_g1) // 96 for p in public_inputs):      # (the MSM truncates to the shorter side)
        raise ValueError('MalformedVerifyingKey')
    if len(public_inputs) != n:
        return False
    transcript.append(b'AB-commitment', proof.com_ab)
    transcript.append(b'C-commitment', proof.com_c)
    if proof.is_lego:
        transcript.append(b'D-commitment', proof.com_d)
    r = transcript.challenge_scalar(b'r-random-fiatshamir')
    checker = pc.RandomizedPairingChecker.new(checker_random, lazy)
    if not verify_tipp_mipp(v, proof, r, transcript, checker):
        return False
    # final_verification_check: prod e(A_i, B_i)^(r^i) == e(alpha^(sum r^i), beta) e(agg inputs, gamma) e(z_c, delta)
    r_powers = powers(r, n)
    r_sum = sum(r_powers) % R_MODULUS
    inp = aggregate_public_inputs(public_inputs, r_powers, r_sum, vk.gamma_abc_g1)
    if clear_d is not None:                                     # using_groth16.rs:106-114: (sum r^i D_i + inputs) pairs with gamma
        d_r = bytes(lib.msm(_cat(clear_d), gp.fr_to_bytes(r_powers)))
        inp = gp.into_affine(gp.add([d_r, gp.to_projective(inp)]))
    source1 = [gp.mul_affine(vk.alpha_g1, r_sum), inp, proof.z_c]
    source2 = [vk.beta_g2, vk.gamma_g2, vk.delta_g2]
    if proof.is_lego:                                           # legogroth16/verifier.rs:92-96
        source1.insert(0, proof.z_d)
        source2.insert(0, vk.gamma_g2)
    checker.add_multiple_sources_and_target(source1, source2, proof.z_ab)
    return checker.verify()
