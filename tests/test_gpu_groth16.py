"""GPU: LegoGroth16 generate -> prove -> verify over the backend (rows a18, a19, f1, f2 of SURVEY.md section 8),
re-expressing legogroth16/src/tests.rs:149-360 (test_prove_and_verify): a proof assembled from the GPU's witness map
and MSMs satisfies verify_qap_proof's pairing equation (verifier.rs:62-84) and verify_witness_commitment
(prover.rs:437-467); wrong public inputs, a tampered proof and a wrong opening are rejected.

Sizes: a 2^10-constraint circuit cross-checked piece by piece against the oracle (h coefficients, every MSM through the
known CRS trapdoor, the pairing equation recomputed by the CPU oracle), and BASELINE config 3's 2^18-constraint circuit
(domain 2^18) end to end."""
import numpy as np
import pytest

from crypto_b200 import groth16 as g16
from crypto_b200 import group as gp
from oracle import bls12_381 as o
from tests import helpers as h
from tools.synth_circuit import synthetic_r1cs

pytestmark = pytest.mark.gpu

R = o.R


def _generators():
    return o.g1_to_bytes(o.G1_GEN), o.g2_to_bytes(o.G2_GEN)


def _setup(num_constraints, cw, seed):
    cs, w = synthetic_r1cs(num_constraints, num_public=2, seed=seed)
    assert cs.is_satisfied(w)
    tox = dict(alpha=0x1111 + seed, beta=0x2222 + seed, gamma=0x3333 + seed, delta=0x4444 + seed, eta=0x5555 + seed)
    t = 0x1234567 + seed
    g1, g2 = _generators()
    pk, ni = g16.generate_parameters(cs, t=t, g1_generator=g1, g2_generator=g2, commit_witness_count=cw, **tox)
    assert ni == cs.num_instance_variables
    return cs, w, pk, tox, t


def _qap_eval_ints(cs, w):
    """A w, B w, C w over the domain as Python ints (r1cs_to_qap.rs:163-186), the oracle side of the witness map."""
    D = g16.domain_size_for(cs.num_constraints + cs.num_instance_variables)
    ev = lambda row: sum(c * w[i] for c, i in row) % R
    a = [ev(r) for r in cs.a] + [0] * (D - cs.num_constraints)
    b = [ev(r) for r in cs.b] + [0] * (D - cs.num_constraints)
    c = [ev(r) for r in cs.c] + [0] * (D - cs.num_constraints)
    a[cs.num_constraints:cs.num_constraints + cs.num_instance_variables] = [x % R for x in w[:cs.num_instance_variables]]
    return a, b, c, D


def test_generate_prove_verify_small_against_the_oracle(dg, cref):
    cw = 3
    cs, w, pk, tox, t = _setup(1 << 10, cw, seed=5)
    ni, nw = cs.num_instance_variables, cs.num_witness_variables
    # -- generator (f2): every query element is its trapdoor scalar times the generator
    a, b, c, zt, qv, m_raw = g16.instance_map_with_evaluation(cs, t)
    assert len(pk.common.a_query) == 96 * (qv + 1) and len(pk.common.h_query) == 96 * (m_raw - 1)
    assert pk.common.a_query == bytes(cref.g1_generator_muls(h.scalars_bytes(a)))
    assert pk.common.b_g2_query == bytes(cref.g2_generator_muls(h.scalars_bytes(b)))
    dinv = pow(tox['delta'], -1, R)
    assert pk.common.h_query == bytes(cref.g1_generator_muls(h.scalars_bytes(g16.h_query_scalars(m_raw - 1, t, zt, dinv))))
    # -- prover (a18 + f1), device-chained
    dpk = g16.DeviceProvingKey(pk, cs)
    r, s, v = 0xabcdef01, 0x1234abcd, 0x777
    proof, committed, hcoef = g16.create_proof(dpk, w, r, s, v, want_h=True)
    assert committed == [x % R for x in w[ni:ni + cw]]
    # h against the oracle's witness-map tail on A w, B w, C w computed with Python integers
    ea, eb, ec, D = _qap_eval_ints(cs, w)
    logD = D.bit_length() - 1
    exp_h = cref.qap_h_from_abc(gp.fr_to_mont(ea), gp.fr_to_mont(eb), gp.fr_to_mont(ec), logD)
    assert bytes(hcoef) == bytes(exp_h)
    # A, B, D through the trapdoor: A = (alpha + sum a_i w_i + r delta) G, B likewise in G2
    aw = sum(x * y for x, y in zip(a, w)) % R
    bw = sum(x * y for x, y in zip(b, w)) % R
    assert proof.a == bytes(cref.g1_generator_muls(h.scalars_bytes([(tox['alpha'] + aw + r * tox['delta']) % R])))
    assert proof.b == bytes(cref.g2_generator_muls(h.scalars_bytes([(tox['beta'] + bw + s * tox['delta']) % R])))
    # -- verifier (a19) on the GPU and, independently, with the CPU oracle's pairing
    pvk = g16.prepare_verifying_key(pk.vk)
    pub = w[1:ni]
    assert g16.verify_proof(pvk, proof, pub)
    assert g16.verify_witness_commitment(pk.vk, proof, len(pub), committed, v)
    d = g16.calculate_d(pvk, proof, pub)
    lhs = cref.multi_pairing(np.frombuffer(proof.a + proof.c + d, dtype=np.uint8),
                             np.frombuffer(proof.b + pvk.delta_g2_neg_pc + pvk.gamma_g2_neg_pc, dtype=np.uint8))
    assert bytes(lhs) == bytes(cref.multi_pairing(np.frombuffer(pk.vk.alpha_g1, dtype=np.uint8), np.frombuffer(pk.vk.beta_g2, dtype=np.uint8)))
    assert bytes(lhs) == pvk.alpha_g1_beta_g2
    # -- rejections: wrong public input, tampered C, wrong opening of D
    assert not g16.verify_proof(pvk, proof, [pub[0], (pub[1] + 1) % R])
    bad = g16.Proof(a=proof.a, b=proof.b, c=gp.mul_affine(proof.c, 2), d=proof.d)
    assert not g16.verify_proof(pvk, bad, pub)
    assert not g16.verify_witness_commitment(pk.vk, proof, len(pub), committed, v + 1)
    assert not g16.verify_witness_commitment(pk.vk, proof, len(pub), [committed[0] + 1] + committed[1:], v)
    # the chained call gives the same proof whatever the number of streams its MSMs are spread over (2 .. 5, default 5)
    try:
        for ns in (2, 3, 4):
            dg.dbg_set_tunable(1, ns)
            p_ns, _ = g16.create_proof(dpk, w, r, s, v)
            assert (p_ns.a, p_ns.b, p_ns.c, p_ns.d) == (proof.a, proof.b, proof.c, proof.d)
    finally:
        dg.dbg_set_tunable(1, 0)
    # r = 0 skips the B-in-G1 MSM (prover.rs:329-339); the proof still verifies
    proof0, _ = g16.create_proof(dpk, w, 0, s, v)
    assert g16.verify_proof(pvk, proof0, pub)
    # a second key object with the same content reuses the resident vectors (content-keyed cache, row f4)
    before = len(g16.KEY_CACHE)
    dpk2 = g16.DeviceProvingKey(pk, cs)
    assert len(g16.KEY_CACHE) == before and dpk2.a_query.handle == dpk.a_query.handle
    dpk2.free()
    dpk.free()
    g16.KEY_CACHE.clear()


def test_unsatisfied_assignment_does_not_verify(dg, cref):
    cs, w, pk, tox, t = _setup(300, 1, seed=9)
    dpk = g16.DeviceProvingKey(pk, cs)
    ni = cs.num_instance_variables
    w_bad = list(w)
    w_bad[ni + 20] = (w_bad[ni + 20] + 1) % R
    assert not cs.is_satisfied(w_bad)
    pvk = g16.prepare_verifying_key(pk.vk)
    good, _ = g16.create_proof(dpk, w, 11, 12, 13)
    bad, _ = g16.create_proof(dpk, w_bad, 11, 12, 13)
    assert g16.verify_proof(pvk, good, w[1:ni])
    assert not g16.verify_proof(pvk, bad, w[1:ni])
    dpk.free()
    g16.KEY_CACHE.clear()


def test_config3_legogroth16_2p18_constraints_end_to_end(dg, cref):
    """BASELINE config 3 at full size: 2^18 - 3 constraints (domain 2^18), ~3.4e5 variables; witness map, four G1 MSMs
    and the G2 MSM chained on the device; the proof verifies on the GPU; h equals the oracle's."""
    cw = 2
    cs, w, pk, tox, t = _setup((1 << 18) - 3, cw, seed=3)
    ni = cs.num_instance_variables
    assert g16.domain_size_for(cs.num_constraints + ni) == 1 << 18
    dpk = g16.DeviceProvingKey(pk, cs, precompute=False)
    proof, committed, hcoef = g16.create_proof(dpk, w, 0x1357, 0x2468, 0x99, want_h=True)
    ea, eb, ec, D = _qap_eval_ints(cs, w)
    exp_h = cref.qap_h_from_abc(gp.fr_to_mont(ea), gp.fr_to_mont(eb), gp.fr_to_mont(ec), 18)
    assert bytes(hcoef) == bytes(exp_h)
    pvk = g16.prepare_verifying_key(pk.vk)
    pub = w[1:ni]
    assert g16.verify_proof(pvk, proof, pub)
    assert g16.verify_witness_commitment(pk.vk, proof, len(pub), committed, 0x99)
    assert not g16.verify_proof(pvk, proof, [pub[0] ^ 1, pub[1]])
    # the resident-table mode of the proving key gives the same proof
    dpk_t = g16.DeviceProvingKey(pk, cs, precompute=True)
    proof_t, _ = g16.create_proof(dpk_t, w, 0x1357, 0x2468, 0x99)
    assert proof_t == proof
    dpk.free()
    dpk_t.free()
    g16.KEY_CACHE.clear()
