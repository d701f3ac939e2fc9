"""GPU: the composite proof of row a20 (proof_system::Proof::new / verify, proof_system/tests/bound_check_legogroth16.rs:
pok_of_bbs_plus_sig_and_bounded_message): knowledge of a BBS+ signature AND a LegoGroth16 proof that one of the signed,
undisclosed messages lies in [min, max], tied together by a witness equality (one shared Schnorr blinding, the response
published once), verified with one challenge and ONE randomized pairing checker (one final exponentiation for the
signature's two pairs and the SNARK's three)."""
import random

import pytest

from crypto_b200 import groth16 as g16
from crypto_b200 import pairing_check as pc
from crypto_b200 import proof_system as ps
from crypto_b200 import protocols as pr
from oracle import bls12_381 as o
from tests import helpers as h

pytestmark = pytest.mark.gpu

R = o.R


def _setup(cref, nmsg, seed):
    rng = random.Random(seed)
    ks = [rng.randrange(1, R) for _ in range(nmsg + 2)]
    pts = bytes(cref.g1_generator_muls(h.scalars_bytes(ks)))
    params = pr.SignatureParamsG1(g1=pts[:96], g2=o.g2_to_bytes(o.G2_GEN), h_0=pts[96:192], h=pts[192:])
    x = rng.randrange(1, R)
    pk = bytes(cref.g2_generator_muls(h.scalars_bytes([x])))
    cs = ps.bound_check_circuit(64)
    snark_pk, ni = g16.generate_parameters(cs, *(rng.randrange(1, R) for _ in range(5)), t=rng.randrange(1 << 100, 1 << 200),
                                           g1_generator=o.g1_to_bytes(o.G1_GEN), g2_generator=o.g2_to_bytes(o.G2_GEN), commit_witness_count=1)
    assert ni == 3 and len(snark_pk.vk.gamma_abc_g1) == 96 * 4
    return rng, params, x, pk, cs, snark_pk


@pytest.mark.parametrize('lazy', [True, False, None])
def test_pok_of_bbs_plus_sig_and_bounded_message(dg, cref, lazy):
    nmsg, bounded = 6, 2
    rng, params, x, pk, cs, snark_pk = _setup(cref, nmsg, 77)
    vmin, vmax = 100, 10_000_000
    msgs = [rng.randrange(R) for _ in range(nmsg)]
    msgs[bounded] = 123_456
    sig = pr.SignatureG1.new(msgs, x, params, rng.randrange(R), rng.randrange(R))
    assert sig.verify(msgs, pk, params)
    dpk = g16.DeviceProvingKey(snark_pk, cs)
    revealed = {0: msgs[0], 5: msgs[5]}
    prover_spec = ps.ProofSpec([ps.PoKBBSSignatureG1(params, pk, revealed), ps.BoundCheckLegoGroth16(vmin, vmax, snark_pk.vk, dpk)],
                               [[(0, bounded), (1, 0)]])
    verifier_spec = ps.ProofSpec([ps.PoKBBSSignatureG1(params, pk, revealed), ps.BoundCheckLegoGroth16(vmin, vmax, snark_pk.vk)],
                                 [[(0, bounded), (1, 0)]])
    rnd = (rng.randrange(1, R) for _ in range(10_000))
    proof = ps.Proof.new(prover_spec, [(sig, msgs), msgs[bounded]], b'nonce-1', rnd)

    def checker():
        return None if lazy is None else pc.RandomizedPairingChecker.new(0x1234567890abcdef1234, lazy)

    assert proof.verify(verifier_spec, checker())
    # a different nonce, other bounds, a wrong revealed message: all rejected
    assert not ps.Proof(proof.statement_proofs, b'nonce-2').verify(verifier_spec, checker())
    tight = ps.ProofSpec([verifier_spec.statements[0], ps.BoundCheckLegoGroth16(vmin, 1000, snark_pk.vk)], verifier_spec.witness_equalities)
    assert not proof.verify(tight, checker())
    wrong = ps.ProofSpec([ps.PoKBBSSignatureG1(params, pk, {0: msgs[0] + 1, 5: msgs[5]}), verifier_spec.statements[1]], verifier_spec.witness_equalities)
    assert not proof.verify(wrong, checker())
    # a prover whose bounded value is NOT the signed message (both proofs fine on their own): the shared response gives it away
    cheat = ps.Proof.new(prover_spec, [(sig, msgs), 5000], b'nonce-1', (rng.randrange(1, R) for _ in range(10_000)))
    assert not cheat.verify(verifier_spec, checker())
    # a signed message outside the bounds: the SNARK's assignment is unsatisfied and the proof does not verify
    msgs2 = list(msgs)
    msgs2[bounded] = 99
    sig2 = pr.SignatureG1.new(msgs2, x, params, rng.randrange(R), rng.randrange(R))
    spec2 = ps.ProofSpec([ps.PoKBBSSignatureG1(params, pk, {0: msgs2[0], 5: msgs2[5]}), prover_spec.statements[1]], prover_spec.witness_equalities)
    out = ps.Proof.new(spec2, [(sig2, msgs2), msgs2[bounded]], b'nonce-1', (rng.randrange(1, R) for _ in range(10_000)))
    vspec2 = ps.ProofSpec([spec2.statements[0], verifier_spec.statements[1]], verifier_spec.witness_equalities)
    assert not out.verify(vspec2, checker())
    dpk.free()
    g16.KEY_CACHE.clear()
