"""GPU: the protocol-level callers of the hot path replayed against the backend at BASELINE config 4 sizes
(SURVEY.md 8a rows a12 - a17, Appendix D):

  Schnorr        schnorr_pok/src/pok_generalized_pedersen.rs:96-161 (tests :300-340): sum r_i G_i - c Y == t
  BBS+ sign      bbs_plus/src/signature.rs:452 signature_verification, 10 000 messages
  BBS+ PoK       bbs_plus/src/proof.rs:1289-1318 pok_signature_revealed_message, test_PoK_multiple_sigs_with_randomized_pairing_check
  accumulator    vb_accumulator/src/batch_utils.rs:716-736 Omega::check, witness.rs:1073-1410 batch witness updates,
                 positive.rs:401-425 membership verification, 10 000 members, |add| = |rem| = 100

Every group element the GPU returns is also compared with its closed form through the known discrete logs
(bases are k_i * G), so the relations are not only self-consistent but bit-exact against the oracle."""
import random

import numpy as np
import pytest

from crypto_b200 import group as gp
from crypto_b200 import pairing_check as pc
from crypto_b200 import protocols as pr
from oracle import bls12_381 as o
from tests import helpers as h

pytestmark = pytest.mark.gpu

R = o.R


def _gmul(cref, k):
    return bytes(cref.g1_generator_muls(h.scalars_bytes([k % R])))


def _params(cref, n, seed):
    """SignatureParamsG1 with known discrete logs: g1, h_0, h_1..h_n = k * G; g2 = the G2 generator."""
    ks = h.ints_of(h.rand_scalars(n + 2, seed))
    pts = bytes(cref.g1_generator_muls(h.scalars_bytes(ks)))
    params = pr.SignatureParamsG1(g1=pts[:96], g2=o.g2_to_bytes(o.G2_GEN), h_0=pts[96:192], h=pts[192:])
    return params, ks


@pytest.mark.parametrize('n', [3, 10000])
def test_schnorr_commitment_and_response(dg, cref, n):
    """SchnorrCommitment::new / response / SchnorrResponse::is_valid with N bases (N = 10^4 + 2 in config 4)."""
    bases, ks = h.g1_bases(n, 50 + n)
    wits = h.ints_of(h.rand_scalars(n, 51 + n))
    blind = h.ints_of(h.rand_scalars(n, 52 + n))
    y = gp.into_affine(pr._msm_unchecked(bytes(bases), wits))
    assert y == h.known_dlog_msm_g1(ks, h.scalars_bytes(wits))
    comm = pr.SchnorrCommitment.new(bytes(bases), blind)
    assert comm.t == h.known_dlog_msm_g1(ks, h.scalars_bytes(blind))
    c = pr.random_oracle_challenge(comm.challenge_contribution(), pr.compressed(y))
    resp = comm.response(wits, c)
    assert resp.is_valid(bytes(bases), y, comm.t, c)
    assert not resp.is_valid(bytes(bases), y, comm.t, c + 1)
    bad = pr.SchnorrResponse(resp.responses[:-1] + [(resp.responses[-1] + 1) % R])
    assert not bad.is_valid(bytes(bases), y, comm.t, c)
    with pytest.raises(ValueError):
        pr.SchnorrResponse(resp.responses[:-1]).is_valid(bytes(bases), y, comm.t, c)


def test_bbs_plus_sign_verify_10k_messages(dg, cref):
    n = 10000
    params, ks = _params(cref, n, 7)
    msgs = h.ints_of(h.rand_scalars(n, 8))
    x, e, s = 0x1234567, 0x7654321, 0x2468ace
    pk = bytes(cref.g2_generator_muls(h.scalars_bytes([x])))
    sig = pr.SignatureG1.new(msgs, x, params, e, s)
    # closed form: A = (k_g1 + k_h0 s + sum k_i m_i) / (e + x) * G
    b_log = (ks[0] + ks[1] * s + sum(k * m for k, m in zip(ks[2:], msgs))) % R
    assert sig.A == _gmul(cref, b_log * pow(e + x, -1, R))
    assert sig.verify(msgs, pk, params)
    tampered = list(msgs)
    tampered[n // 2] += 1
    assert not sig.verify(tampered, pk, params)
    assert not pr.SignatureG1(sig.A, sig.e + 1, sig.s).verify(msgs, pk, params)


def test_bbs_plus_pok_of_signature_10k_messages(dg, cref):
    """pok_signature_revealed_message at config 4 size: two ~10^4-term MSMs on the prover side, two on the verifier side
    plus the 2-pair check."""
    n = 10000
    params, ks = _params(cref, n, 17)
    msgs = h.ints_of(h.rand_scalars(n, 18))
    x, e, s = 0x1111, 0x2222, 0x3333
    pk = bytes(cref.g2_generator_muls(h.scalars_bytes([x])))
    sig = pr.SignatureG1.new(msgs, x, params, e, s)
    revealed = {0: msgs[0], 17: msgs[17], n - 1: msgs[n - 1]}
    rnd = h.ints_of(h.rand_scalars(n + 8, 19))
    proto = pr.PoKOfSignatureG1Protocol.init(sig, params, msgs, set(revealed), rnd)
    # closed forms of the randomised signature
    r1, r2 = rnd[0], rnd[1]
    b_log = (ks[0] + ks[1] * s + sum(k * m for k, m in zip(ks[2:], msgs))) % R
    a_log = b_log * pow(e + x, -1, R) % R
    assert proto.A_prime == _gmul(cref, a_log * r1)
    assert proto.A_bar == _gmul(cref, r1 * b_log - e * a_log * r1)
    assert proto.d == _gmul(cref, r1 * b_log - r2 * ks[1])
    c = pr.random_oracle_challenge(proto.challenge_contribution(revealed, params))
    proof = proto.gen_proof(c)
    assert proof.verify(revealed, c, pk, params)
    assert not proof.verify(revealed, c + 1, pk, params)
    wrong = dict(revealed)
    wrong[17] += 1
    assert not proof.verify(wrong, c, pk, params)
    other_pk = bytes(cref.g2_generator_muls(h.scalars_bytes([x + 1])))
    assert not proof.verify(revealed, c, other_pk, params)


@pytest.mark.parametrize('lazy', [True, False])
def test_pok_multiple_sigs_with_randomized_pairing_check(dg, cref, lazy):
    """test_PoK_multiple_sigs_with_randomized_pairing_check with 100 signatures: 100 x (A', A_bar) pairs -> 200 Miller
    loops and ONE final exponentiation in the lazy checker; a single bad proof makes the whole batch fail."""
    nsig, n = 100, 5
    params, ks = _params(cref, n, 23)
    x = 0xabcdef
    pk = bytes(cref.g2_generator_muls(h.scalars_bytes([x])))
    rng = random.Random(5)
    proofs = []
    for i in range(nsig):
        msgs = [rng.randrange(R) for _ in range(n)]
        sig = pr.SignatureG1.new(msgs, x, params, rng.randrange(R), rng.randrange(R))
        revealed = {1: msgs[1]}
        proto = pr.PoKOfSignatureG1Protocol.init(sig, params, msgs, set(revealed), [rng.randrange(1, R) for _ in range(n + 8)])
        c = pr.random_oracle_challenge(proto.challenge_contribution(revealed, params))
        proofs.append((proto.gen_proof(c), revealed, c))
    ck = pc.RandomizedPairingChecker.new(0x1234567890abcdef, lazy)
    for proof, revealed, c in proofs:
        assert proof.verify_with_randomized_pairing_checker(revealed, c, pk, params, ck)
    assert ck.verify()
    # individually verified proofs agree with the batch
    assert proofs[0][0].verify(proofs[0][1], proofs[0][2], pk, params)
    bad = pc.RandomizedPairingChecker.new(0x1234567890abcdef, lazy)
    for i, (proof, revealed, c) in enumerate(proofs):
        if i == 57:
            proof = pr.PoKOfSignatureG1Proof(proof.A_prime, gp.mul_affine(proof.A_bar, 2), proof.d, proof.sc_resp_1, proof.T2, proof.sc_resp_2)
        bad.add_sources(proof.A_prime, pk, proof.A_bar, params.g2)
    assert not bad.verify()


def _accumulator(cref, nmem, nadd, nrem, seed):
    rng = random.Random(seed)
    alpha = rng.randrange(1, R)
    members = [rng.randrange(R) for _ in range(nmem)]
    additions = [rng.randrange(R) for _ in range(nadd)]
    removals = members[:nrem]
    u = 1
    for m in members:
        u = u * (m + alpha) % R
    V = _gmul(cref, u)
    return alpha, members, additions, removals, u, V


@pytest.mark.parametrize('nadd,nrem', [(100, 100), (30, 0), (0, 20), (1, 1)])
def test_omega_check(dg, cref, nadd, nrem):
    """Omega::check: <powers of y, Omega> / d_D(y) == V * v_AD(y) / d_D(y); Omega::new's coefficients * V by closed form."""
    alpha, members, additions, removals, u, V = _accumulator(cref, 200, nadd, nrem, 3 + nadd)
    omega = pr.Omega.new(additions, removals, V, alpha)
    coeffs = pr.Poly_v_AD.generate(additions, removals, alpha)
    assert len(omega) == len(coeffs)
    assert omega.points == bytes(cref.g1_generator_muls(h.scalars_bytes([c * u % R for c in coeffs])))
    for y in (members[-1], members[-2], 12345):
        v_AD = pr.Poly_v_AD.eval_direct(additions, removals, alpha, y)
        d_D_inv = pow(pr.Poly_d.eval_direct(removals, y), -1, R)
        V_prime = gp.mul_affine(V, v_AD * d_D_inv)
        assert gp.into_affine(omega.evaluate(y, d_D_inv)) == V_prime == _gmul(cref, u * v_AD * d_D_inv)


def test_accumulator_batch_witness_update_10k_members(dg, cref):
    """BASELINE config 4: 10 000 witnesses updated after |add| = |rem| = 100 with the secret key (one fused device call);
    every new witness equals V' / (y + alpha) by closed form, sampled ones pass the pairing membership check against V',
    and the Omega-based public update gives the same witness."""
    nmem, nadd, nrem = 10100, 100, 100
    alpha, members, additions, removals, u, V = _accumulator(cref, nmem, nadd, nrem, 11)
    elements = members[nrem:]                                     # the 10 000 members that stay
    assert len(elements) == 10000
    wit_logs = [u * pow(y + alpha, -1, R) % R for y in elements]
    old_wits = bytes(cref.g1_generator_muls(h.scalars_bytes(wit_logs)))
    d_factors, new_wits = pr.compute_update_using_secret_key_after_batch_updates(additions, removals, elements, old_wits, V, alpha)
    u_new = u
    for a in additions:
        u_new = u_new * (a + alpha) % R
    for d in removals:
        u_new = u_new * pow(d + alpha, -1, R) % R
    V_new = _gmul(cref, u_new)
    exp = bytes(cref.g1_generator_muls(h.scalars_bytes([u_new * pow(y + alpha, -1, R) % R for y in elements])))
    assert new_wits == exp
    P_tilde = o.g2_to_bytes(o.G2_GEN)
    Q_tilde = bytes(cref.g2_generator_muls(h.scalars_bytes([alpha])))
    for i in (0, 4999, 9999):
        w = new_wits[96 * i:96 * i + 96]
        assert pr.verify_membership_given_accumulated(V_new, elements[i], w, Q_tilde, P_tilde)
        assert not pr.verify_membership_given_accumulated(V, elements[i], w, Q_tilde, P_tilde)
        assert not pr.verify_membership_given_accumulated(V_new, elements[i] + 1, w, Q_tilde, P_tilde)
    omega = pr.Omega.new(additions, removals, V, alpha)
    for i in (1, 7777):
        w = pr.compute_update_using_public_info_after_batch_updates(additions, removals, omega, elements[i], old_wits[96 * i:96 * i + 96])
        assert w == new_wits[96 * i:96 * i + 96]
    with pytest.raises(ValueError):
        pr.compute_update_using_secret_key_after_batch_updates(additions, removals, elements[:-1], old_wits, V, alpha)
