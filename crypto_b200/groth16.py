"""Host-side mirror of the LegoGroth16 calls that sit on the hot path, over the C ABI (BLS12-381 only).

Same names, argument meaning and control flow as
  * legogroth16/src/r1cs_to_qap.rs:100-258    LibsnarkReduction::{instance_map_with_evaluation, h_query_scalars};
                                               witness_map_from_matrices runs on the device (dg_groth16_prove_msms)
  * legogroth16/src/generator.rs:245-442      generate_parameters_and_extra_info_with_qap (toxic waste given explicitly)
  * legogroth16/src/prover.rs:267-383,585-594 create_proof_and_committed_witnesses_with_assignment, calculate_coeff
  * legogroth16/src/prover.rs:437-467         verify_witness_commitment
  * legogroth16/src/verifier.rs:18-108        prepare_verifying_key, prepare_inputs, calculate_d, verify_qap_proof, verify_proof
  * legogroth16/src/data_structures.rs:7-189  VerifyingKey, ProvingKeyCommon, ProvingKey, Proof
so the parity tests read like legogroth16/src/tests.rs.  Every group operation (fixed-base tables, MSMs, scalar
multiplications, sums, normalisation, Miller loops, final exponentiation) runs on the GPU; the host does the Fr
bookkeeping with Python integers, which is what stays on the CPU in the Rust glue as well.

Group elements are packed Montgomery records (bytes): affine G1 96 B, affine G2 192 B, GT 576 B.
"""
import hashlib
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import group as gp
from . import lib
from .group import G1, G2, R_MODULUS


# ---- constraint system ------------------------------------------------------------------------------------------
@dataclass
class ConstraintMatrices:
    """ark_relations::r1cs::ConstraintMatrices: rows of (coefficient, variable index); variable 0 is the constant 1,
    instance variables come first, then the witness variables."""
    num_instance_variables: int
    num_witness_variables: int
    a: List[list] = field(default_factory=list)
    b: List[list] = field(default_factory=list)
    c: List[list] = field(default_factory=list)

    @property
    def num_constraints(self):
        return len(self.a)

    @property
    def num_variables(self):
        return self.num_instance_variables + self.num_witness_variables

    def csr(self):
        """Three (row_ptr, col, coeff_mont) triples in the layout dg_r1cs_upload / dg_fr_spmv take."""
        out = []
        for mat in (self.a, self.b, self.c):
            rp = np.zeros(len(mat) + 1, dtype=np.uint32)
            cols, coeffs = [], []
            for i, row in enumerate(mat):
                for coeff, idx in row:
                    cols.append(idx)
                    coeffs.append(coeff)
                rp[i + 1] = len(cols)
            out.append((rp, np.asarray(cols, dtype=np.uint32), gp.fr_to_mont(coeffs)))
        return out

    def is_satisfied(self, full_assignment):
        def ev(row):
            return sum(c * full_assignment[i] for c, i in row) % R_MODULUS
        return all(ev(a) * ev(b) % R_MODULUS == ev(c) for a, b, c in zip(self.a, self.b, self.c))


# ---- evaluation domain (ark_poly Radix2EvaluationDomain over Fr) --------------------------------------------------
def domain_size_for(n):
    return 1 << max((n - 1).bit_length(), 0)


def domain_generator(size):
    """F::get_root_of_unity(size): the 2^32-th root 7^((r-1)/2^32) squared down to order `size`."""
    return pow(7, (R_MODULUS - 1) // size, R_MODULUS)


def batch_inverse(vals):
    """Montgomery's trick (ark_ff::batch_inversion); every value must be non-zero."""
    pref, acc = [], 1
    for v in vals:
        pref.append(acc)
        acc = acc * v % R_MODULUS
    inv = pow(acc, -1, R_MODULUS)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * pref[i] % R_MODULUS
        inv = inv * vals[i] % R_MODULUS
    return out


def evaluate_all_lagrange_coefficients(size, t):
    """EvaluationDomain::evaluate_all_lagrange_coefficients for t outside the domain: u_i = Z(t) w^i / (n (t - w^i))."""
    w = domain_generator(size)
    zt = (pow(t, size, R_MODULUS) - 1) % R_MODULUS
    if zt == 0:
        raise ValueError('t lies in the domain')
    pows, cur = [], 1
    for _ in range(size):
        pows.append(cur)
        cur = cur * w % R_MODULUS
    inv = batch_inverse([(t - x) % R_MODULUS for x in pows])
    k = zt * pow(size, -1, R_MODULUS) % R_MODULUS
    return [k * x % R_MODULUS * iv % R_MODULUS for x, iv in zip(pows, inv)], zt


def instance_map_with_evaluation(cs: ConstraintMatrices, t):
    """LibsnarkReduction::instance_map_with_evaluation (r1cs_to_qap.rs:104-148) -> (a, b, c, zt, qap_num_variables, domain_size)."""
    domain_size = domain_size_for(cs.num_constraints + cs.num_instance_variables)
    u, zt = evaluate_all_lagrange_coefficients(domain_size, t)
    qap_num_variables = (cs.num_instance_variables - 1) + cs.num_witness_variables
    a = [0] * (qap_num_variables + 1)
    b = [0] * (qap_num_variables + 1)
    c = [0] * (qap_num_variables + 1)
    nc = cs.num_constraints
    a[0:cs.num_instance_variables] = u[nc:nc + cs.num_instance_variables]
    for i in range(nc):
        u_i = u[i]
        for coeff, index in cs.a[i]:
            a[index] = (a[index] + u_i * coeff) % R_MODULUS
        for coeff, index in cs.b[i]:
            b[index] = (b[index] + u_i * coeff) % R_MODULUS
        for coeff, index in cs.c[i]:
            c[index] = (c[index] + u_i * coeff) % R_MODULUS
    return a, b, c, zt, qap_num_variables, domain_size


def h_query_scalars(max_power, t, zt, delta_inverse):
    k = zt * delta_inverse % R_MODULUS
    out, cur = [], 1
    for _ in range(max_power):
        out.append(k * cur % R_MODULUS)
        cur = cur * t % R_MODULUS
    return out


# ---- keys and proofs (data_structures.rs) -------------------------------------------------------------------------
@dataclass
class VerifyingKey:
    alpha_g1: bytes
    beta_g2: bytes
    gamma_g2: bytes
    delta_g2: bytes
    gamma_abc_g1: bytes          # concatenated affine records
    eta_gamma_inv_g1: bytes
    commit_witness_count: int

    def num_public_inputs(self):
        return len(self.gamma_abc_g1) // 96 - self.commit_witness_count


@dataclass
class ProvingKeyCommon:
    beta_g1: bytes
    delta_g1: bytes
    eta_delta_inv_g1: bytes
    a_query: bytes
    b_g1_query: bytes
    b_g2_query: bytes
    h_query: bytes
    l_query: bytes


@dataclass
class ProvingKey:
    vk: VerifyingKey
    common: ProvingKeyCommon


@dataclass
class Proof:
    a: bytes
    b: bytes
    c: bytes
    d: bytes


@dataclass
class PreparedVerifyingKey:
    vk: VerifyingKey
    alpha_g1_beta_g2: bytes
    gamma_g2_neg_pc: bytes
    delta_g2_neg_pc: bytes


def _fixed_base_msm(generator, window_hint, scalars, group):
    """FixedBase::get_window_table + FixedBase::msm + normalize_batch (generator.rs:335-425) on the device."""
    tbl = lib.FixedBaseTable(generator, window_hint, g2=group.g2)
    try:
        return bytes(tbl.mul_many_normalized(gp.fr_to_bytes(scalars)))
    finally:
        tbl.free()


def generate_parameters(cs: ConstraintMatrices, alpha, beta, gamma, delta, eta, t, g1_generator, g2_generator,
                        commit_witness_count):
    """generate_parameters_and_extra_info_with_qap with the toxic waste and the evaluation point t supplied by the
    caller (the reference samples them from its rng)."""
    num_instance_variables = cs.num_instance_variables
    if cs.num_witness_variables < commit_witness_count:
        raise ValueError('InsufficientWitnessesForCommitment(%d, %d)' % (cs.num_witness_variables, commit_witness_count))
    n = num_instance_variables + commit_witness_count
    a, b, c, zt, qap_num_variables, m_raw = instance_map_with_evaluation(cs, t)
    non_zero_a = sum(1 for x in a if x)
    non_zero_b = sum(1 for x in b if x)
    gamma_inverse = pow(gamma, -1, R_MODULUS)
    delta_inverse = pow(delta, -1, R_MODULUS)
    lin = [(beta * x + alpha * y + z) % R_MODULUS for x, y, z in zip(a, b, c)]
    gamma_abc = [v * gamma_inverse % R_MODULUS for v in lin[:n]]
    l = [v * delta_inverse % R_MODULUS for v in lin]
    b_g2_query = _fixed_base_msm(g2_generator, non_zero_b, b, G2)
    g1_hint = non_zero_a + non_zero_b + qap_num_variables + m_raw + 1
    g1_tbl = lib.FixedBaseTable(g1_generator, g1_hint)
    try:
        def g1_msm(scalars):
            return bytes(g1_tbl.mul_many_normalized(gp.fr_to_bytes(scalars))) if len(scalars) else b''
        a_query = g1_msm(a)
        b_g1_query = g1_msm(b)
        h_query = g1_msm(h_query_scalars(m_raw - 1, t, zt, delta_inverse))
        l_query = g1_msm(l[n:])
        gamma_abc_g1 = g1_msm(gamma_abc)
    finally:
        g1_tbl.free()
    vk = VerifyingKey(
        alpha_g1=gp.mul_affine(g1_generator, alpha), beta_g2=gp.mul_affine(g2_generator, beta, G2),
        gamma_g2=gp.mul_affine(g2_generator, gamma, G2), delta_g2=gp.mul_affine(g2_generator, delta, G2),
        gamma_abc_g1=gamma_abc_g1, eta_gamma_inv_g1=gp.mul_affine(g1_generator, eta * gamma_inverse),
        commit_witness_count=commit_witness_count)
    common = ProvingKeyCommon(
        beta_g1=gp.mul_affine(g1_generator, beta), delta_g1=gp.mul_affine(g1_generator, delta),
        eta_delta_inv_g1=gp.mul_affine(g1_generator, eta * delta_inverse),
        a_query=a_query, b_g1_query=b_g1_query, b_g2_query=b_g2_query, h_query=h_query, l_query=l_query)
    return ProvingKey(vk, common), num_instance_variables


# ---- device-resident proving key, keyed by content (data_structures.rs:151-168) ------------------
class DeviceKeyCache:
    """Bases handles keyed by a digest of the affine records: loading the same ProvingKeyCommon twice (two provers of one
    circuit, a key deserialised again) reuses the resident vectors instead of uploading 100s of MB again."""

    def __init__(self):
        self._h = {}

    @staticmethod
    def digest(points, g2, precompute):
        hsh = hashlib.blake2b(digest_size=16)
        hsh.update(b'G2' if g2 else b'G1')
        hsh.update(bytes([1 if precompute else 0]))
        hsh.update(memoryview(np.ascontiguousarray(np.frombuffer(points, dtype=np.uint8))))
        return hsh.digest()

    def get(self, points, g2=False, precompute=False):
        key = self.digest(points, g2, precompute)
        ent = self._h.get(key)
        if ent is None:
            ent = lib.Bases(np.frombuffer(points, dtype=np.uint8), g2=g2)
            if precompute:
                ent.precompute(0)
            self._h[key] = ent
        return ent

    def __len__(self):
        return len(self._h)

    def clear(self):
        for ent in self._h.values():
            ent.free()
        self._h.clear()


KEY_CACHE = DeviceKeyCache()


class DeviceProvingKey:
    """A ProvingKey plus its circuit resident on the GPU: the five query vectors behind bases handles (from the content
    cache), the committed-witness slice of gamma_abc_g1, and the constraint matrices (dg_r1cs_upload)."""

    def __init__(self, pk: ProvingKey, cs: ConstraintMatrices, precompute=False, cache=KEY_CACHE):
        self.pk, self.cs = pk, cs
        c, vk = pk.common, pk.vk
        self.h_query = cache.get(c.h_query, precompute=precompute)
        self.l_query = cache.get(c.l_query, precompute=precompute) if c.l_query else None
        self.a_query = cache.get(c.a_query, precompute=precompute)
        self.b_g1_query = cache.get(c.b_g1_query, precompute=precompute)
        self.b_g2_query = cache.get(c.b_g2_query, g2=True, precompute=precompute)
        ni, cw = cs.num_instance_variables, vk.commit_witness_count
        self.gamma_abc_committed = cache.get(vk.gamma_abc_g1[96 * ni:96 * (ni + cw)]) if cw else None
        self.r1cs = lib.R1CS(cs.csr(), cs.num_constraints, cs.num_instance_variables, cs.num_variables)

    def free(self):
        self.r1cs.free()


def create_proof(dpk: DeviceProvingKey, full_assignment, r, s, v, want_h=False):
    """create_proof_and_committed_witnesses_with_assignment (prover.rs:267-383) with the witness map and all MSMs chained
    on the device.  full_assignment = instance assignment (starting with 1) followed by the witness assignment, as ints.
    -> (Proof, committed_witnesses[, h])"""
    pk, cs = dpk.pk, dpk.cs
    vk, common = pk.vk, pk.common
    ni, nw, cw = cs.num_instance_variables, cs.num_witness_variables, vk.commit_witness_count
    if len(full_assignment) != ni + nw or full_assignment[0] % R_MODULUS != 1:
        raise ValueError('full_assignment must hold 1, the instance variables and the witness variables')
    # calculate_coeff(initial, query, vk_param, assignment) = initial + query[0] + msm(query[1..], assignment) + vk_param;
    # the assignment starts with the constant 1, so ONE msm over the whole query with the full assignment gives
    # query[0] + msm(query[1..], assignment) as the same group element.
    jobs, names = [], []

    def job(name, bases, off, cnt):
        if bases is not None and cnt:
            names.append(name)
            jobs.append((bases, off, cnt))

    job('l', dpk.l_query, ni + cw, nw - cw)
    job('a', dpk.a_query, 0, ni + nw)
    if r % R_MODULUS:
        job('b_g1', dpk.b_g1_query, 0, ni + nw)
    job('b_g2', dpk.b_g2_query, 0, ni + nw)
    job('gamma_abc', dpk.gamma_abc_committed, ni, cw)
    h_acc, res, h = lib.groth16_prove_msms(dpk.r1cs, gp.fr_to_mont(full_assignment), dpk.h_query, jobs, want_h=want_h)
    acc = {nm: bytes(x) for nm, x in zip(names, res)}
    zero1, zero2 = gp.to_projective(bytes(96)), gp.to_projective(bytes(192), G2)
    l_aux_acc = acc.get('l', zero1)
    v_eta_delta_inv = gp.mul(common.eta_delta_inv_g1, v)
    # window-3 table of delta_g1 for the three small multiplications (prover.rs:309-313)
    delta_tbl = lib.FixedBaseTable(common.delta_g1, 3)
    try:
        r_g1, s_g1, rs_g1 = gp.split(delta_tbl.mul_many(gp.fr_to_bytes([r, s, r * s])), 144)
    finally:
        delta_tbl.free()
    g_a = gp.add([r_g1, acc['a'], gp.to_projective(vk.alpha_g1)])
    g1_b = gp.add([s_g1, acc['b_g1'], gp.to_projective(common.beta_g1)]) if r % R_MODULUS else zero1
    s_g2 = gp.mul(vk.delta_g2, s, G2)
    g2_b = gp.add([s_g2, acc['b_g2'], gp.to_projective(vk.beta_g2, G2)], G2)
    g_a_aff, g1_b_aff = gp.into_affine(g_a), gp.into_affine(g1_b)
    g_c = gp.add([gp.mul(g_a_aff, s), gp.mul(g1_b_aff, r), gp.to_projective(gp.neg(gp.into_affine(rs_g1))), l_aux_acc, bytes(h_acc),
                  gp.to_projective(gp.neg(gp.into_affine(v_eta_delta_inv)))])
    g_d = gp.add([acc.get('gamma_abc', zero1), gp.mul(vk.eta_gamma_inv_g1, v)])
    proof = Proof(a=g_a_aff, b=gp.into_affine(g2_b, G2), c=gp.into_affine(g_c), d=gp.into_affine(g_d))
    committed = [x % R_MODULUS for x in full_assignment[ni:ni + cw]]
    return (proof, committed, h) if want_h else (proof, committed)


def verify_witness_commitment(vk: VerifyingKey, proof: Proof, public_inputs_count, witnesses_expected_in_commitment, v):
    """prover.rs:437-467: proof.d == msm(gamma_abc_g1[1 + pub .. ], committed) + eta_gamma_inv_g1 * v."""
    k = len(witnesses_expected_in_commitment)
    if public_inputs_count + k + 1 > len(vk.gamma_abc_g1) // 96:
        raise ValueError('VectorLongerThanExpected')
    lo = 1 + public_inputs_count
    d = bytes(lib.msm(np.frombuffer(vk.gamma_abc_g1[96 * lo:96 * (lo + k)], dtype=np.uint8), gp.fr_to_bytes(witnesses_expected_in_commitment)))
    d = gp.add([d, gp.mul(vk.eta_gamma_inv_g1, v)])
    return gp.into_affine(d) == proof.d


# ---- verifier (verifier.rs) ---------------------------------------------------------------------------------------
def prepare_verifying_key(vk: VerifyingKey):
    return PreparedVerifyingKey(vk=vk, alpha_g1_beta_g2=bytes(lib.multi_pairing(vk.alpha_g1, vk.beta_g2)),
                                gamma_g2_neg_pc=gp.neg(vk.gamma_g2, G2), delta_g2_neg_pc=gp.neg(vk.delta_g2, G2))


def prepare_inputs(pvk: PreparedVerifyingKey, public_inputs):
    gabc = pvk.vk.gamma_abc_g1
    if len(public_inputs) + 1 > len(gabc) // 96:
        raise ValueError('MalformedVerifyingKey')
    if len(public_inputs) > 2:
        inp = [1] + list(public_inputs)
        return bytes(lib.msm(np.frombuffer(gabc, dtype=np.uint8), gp.fr_to_bytes(inp)))
    parts = [gp.to_projective(gabc[:96])]
    for i, x in enumerate(public_inputs):
        parts.append(gp.mul(gabc[96 * (i + 1):96 * (i + 2)], x))
    return gp.add(parts)


def calculate_d(pvk, proof: Proof, public_inputs):
    return gp.into_affine(gp.add([prepare_inputs(pvk, public_inputs), gp.to_projective(proof.d)]))


def verify_qap_proof(pvk: PreparedVerifyingKey, a, b, c, d):
    """e(a, b) e(c, -delta) e(d, -gamma) == e(alpha, beta): one 3-pair Miller loop + one final exponentiation."""
    qap = lib.multi_miller_loop(a + c + d, b + pvk.delta_g2_neg_pc + pvk.gamma_g2_neg_pc)
    fe = lib.final_exponentiation(qap)
    if fe is None:
        raise ValueError('UnexpectedIdentity')
    return bytes(fe) == pvk.alpha_g1_beta_g2


def verify_proof(pvk: PreparedVerifyingKey, proof: Proof, public_inputs):
    return verify_qap_proof(pvk, proof.a, proof.b, proof.c, calculate_d(pvk, proof, public_inputs))
