"""Host-side mirror of the composite-proof orchestration for the statements the hot path serves:

  * proof_system/src/prover.rs:139-...   Proof::new: one Fiat-Shamir challenge over every sub-protocol's contribution,
                                         shared blindings for witnesses declared equal (MetaStatement::WitnessEquality)
  * proof_system/src/verifier.rs:128-150, 1830-1831   Proof::verify with one RandomizedPairingChecker for all statements
  * proof_system/src/sub_protocols/bbs_plus.rs        PoKBBSSignatureG1
  * proof_system/src/sub_protocols/bound_check_legogroth16.rs:84-290   BoundCheckLegoGroth16: a LegoGroth16 proof that a
        message lies in [min, max] plus a Schnorr proof of knowledge of the opening of the proof's commitment D, whose
        response for the message is NOT sent: the verifier takes it from the signature's proof, which is what ties the two
  * proof_system/src/sub_protocols/schnorr.rs         SchnorrProtocol with partial responses

The orchestration itself is host logic (maps, hashing, Fr arithmetic); every group operation it triggers -- the BBS+ MSMs,
the LegoGroth16 witness map and MSMs, the Schnorr commitments, the Miller loops and the ONE final exponentiation of the
checker -- runs on the GPU through crypto_b200.protocols / crypto_b200.groth16 / crypto_b200.pairing_check.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

from . import group as gp
from . import groth16 as g16
from . import protocols as pr
from .group import R_MODULUS


# ---- the bound-check circuit --------------------------------------------------------------------------------------
def bound_check_circuit(nbits=64):
    """min <= value <= max for nbits-bit gaps.  Variables: 0 one, 1 min, 2 max (instance); 3 value (the committed
    witness), then the bits of value - min and of max - value.  Same statement as proof_system's BoundCheckCircuit
    (public inputs min and max, one committed witness); the constraint system is this repo's own."""
    cs = g16.ConstraintMatrices(num_instance_variables=3, num_witness_variables=1 + 2 * nbits)
    lo0, hi0 = 4, 4 + nbits
    for base in (lo0, hi0):
        for i in range(nbits):                              # b (1 - b) = 0
            cs.a.append([(1, base + i)])
            cs.b.append([(1, 0), (R_MODULUS - 1, base + i)])
            cs.c.append([])
    cs.a.append([(1 << i, lo0 + i) for i in range(nbits)])  # sum 2^i b_i = value - min
    cs.b.append([(1, 0)])
    cs.c.append([(1, 3), (R_MODULUS - 1, 1)])
    cs.a.append([(1 << i, hi0 + i) for i in range(nbits)])  # sum 2^i c_i = max - value
    cs.b.append([(1, 0)])
    cs.c.append([(1, 2), (R_MODULUS - 1, 3)])
    return cs


def bound_check_assignment(value, vmin, vmax, nbits=64):
    lo, hi = (value - vmin) % R_MODULUS, (vmax - value) % R_MODULUS
    bits = lambda x: [(x >> i) & 1 for i in range(nbits)]        # an out-of-range gap does not fit: the proof will not verify
    return [1, vmin, vmax, value] + bits(lo) + bits(hi)


# ---- Schnorr protocol with partial responses (sub_protocols/schnorr.rs) ---------------------------------------------
class SchnorrProtocol:
    def __init__(self, commitment_key: List[bytes], commitment: bytes):
        self.key, self.commitment = [bytes(k) for k in commitment_key], bytes(commitment)
        self.comm = None
        self.witnesses = None

    def init(self, blindings: List[int], witnesses: List[int]):
        self.comm = pr.SchnorrCommitment.new(self.key, blindings)
        self.witnesses = list(witnesses)

    def challenge_contribution(self):
        return pr.compressed(b''.join(self.key) + self.commitment + self.comm.t)

    def gen_partial_proof(self, challenge, skip_for):
        resp = self.comm.response(self.witnesses, challenge).responses
        return {'t': self.comm.t, 'responses': {i: r for i, r in enumerate(resp) if i not in skip_for}, 'n': len(resp)}

    def verify_partial_proof(self, challenge, proof, missing: Dict[int, int]):
        full = dict(proof['responses'])
        full.update(missing)
        if sorted(full) != list(range(proof['n'])):
            return False
        return pr.SchnorrResponse([full[i] for i in range(proof['n'])]).is_valid(self.key, self.commitment, proof['t'], challenge)


# ---- statements -------------------------------------------------------------------------------------------------------
@dataclass
class PoKBBSSignatureG1:
    params: pr.SignatureParamsG1
    public_key: bytes
    revealed_messages: Dict[int, int]


@dataclass
class BoundCheckLegoGroth16:
    vmin: int
    vmax: int
    verifying_key: g16.VerifyingKey
    proving_key: Optional[g16.DeviceProvingKey] = None        # prover side only

    @staticmethod
    def schnorr_comm_key(vk: g16.VerifyingKey):
        return [vk.gamma_abc_g1[96 * 3:96 * 4], vk.eta_gamma_inv_g1]


@dataclass
class ProofSpec:
    statements: list
    witness_equalities: List[List[Tuple[int, int]]]           # each: [(statement id, witness id), ...] proven equal


@dataclass
class Proof:
    statement_proofs: list
    nonce: bytes

    @classmethod
    def new(cls, spec: ProofSpec, witnesses: list, nonce: bytes, rnd):
        """witnesses[i]: (SignatureG1, messages) for PoKBBSSignatureG1, the bounded message for BoundCheckLegoGroth16.
        rnd: iterator of random scalars (the prover's rng)."""
        rnd = iter(rnd)
        if len(spec.statements) != len(witnesses):
            raise ValueError('UnequalWitnessAndStatementCount')
        blindings = {}
        for eq in spec.witness_equalities:
            b = next(rnd)
            for ref in eq:
                blindings[tuple(ref)] = b
        protos, contrib = [], []
        for sid, (st, wit) in enumerate(zip(spec.statements, witnesses)):
            if isinstance(st, PoKBBSSignatureG1):
                sig, messages = wit
                own = {j: b for (s, j), b in blindings.items() if s == sid}
                p = pr.PoKOfSignatureG1Protocol(sig, st.params, messages, set(st.revealed_messages),
                                                [next(rnd) for _ in range(len(messages) + 8)], own)
                protos.append(p)
                contrib.append(p.challenge_contribution(st.revealed_messages, st.params))
            elif isinstance(st, BoundCheckLegoGroth16):
                v, r, s = next(rnd), next(rnd), next(rnd)
                snark_proof, _ = g16.create_proof(st.proving_key, bound_check_assignment(wit, st.vmin, st.vmax), r, s, v)
                sp = SchnorrProtocol(st.schnorr_comm_key(st.verifying_key), snark_proof.d)
                sp.init([blindings.get((sid, 0)) or next(rnd), next(rnd)], [wit, v])
                protos.append((snark_proof, sp))
                contrib.append(sp.challenge_contribution())
            else:
                raise ValueError('unsupported statement')
        challenge = pr.random_oracle_challenge(*contrib, nonce)
        out = []
        for st, p in zip(spec.statements, protos):
            if isinstance(st, PoKBBSSignatureG1):
                out.append(p.gen_proof(challenge))
            else:
                snark_proof, sp = p
                out.append({'snark_proof': snark_proof, 'sp': sp.gen_partial_proof(challenge, {0})})
        return cls(out, nonce)

    def verify(self, spec: ProofSpec, pairing_checker=None):
        """One challenge, one (optional) randomized pairing checker for every statement; witness equalities are enforced
        by taking the missing Schnorr response of one statement from the other's proof."""
        if len(spec.statements) != len(self.statement_proofs):
            return False
        contrib = []
        for st, sp in zip(spec.statements, self.statement_proofs):
            if isinstance(st, PoKBBSSignatureG1):
                t1 = sp.sc_resp_1.t
                out = pr.compressed(sp.A_prime + sp.A_bar + sp.d + t1 + sp.T2)
                for i in sorted(st.revealed_messages):
                    out += i.to_bytes(8, 'little') + (st.revealed_messages[i] % R_MODULUS).to_bytes(32, 'little')
                contrib.append(out)
            else:
                key = st.schnorr_comm_key(st.verifying_key)
                contrib.append(pr.compressed(b''.join(key) + sp['snark_proof'].d + sp['sp']['t']))
        challenge = pr.random_oracle_challenge(*contrib, self.nonce)
        # responses the equalities tie together: (statement, witness) -> response published by a signature proof
        shared = {}
        for eq in spec.witness_equalities:
            resp = None
            for sid, wid in eq:
                st = spec.statements[sid]
                if isinstance(st, PoKBBSSignatureG1):
                    r = self.statement_proofs[sid].get_resp_for_message(wid, set(st.revealed_messages))
                    if resp is not None and resp != r:
                        return False                         # two signatures disagree on an "equal" message
                    resp = r
            for ref in eq:
                shared[tuple(ref)] = resp
        for sid, (st, sp) in enumerate(zip(spec.statements, self.statement_proofs)):
            if isinstance(st, PoKBBSSignatureG1):
                if pairing_checker is not None:
                    ok = sp.verify_with_randomized_pairing_checker(st.revealed_messages, challenge, st.public_key, st.params, pairing_checker)
                else:
                    ok = sp.verify(st.revealed_messages, challenge, st.public_key, st.params)
                if not ok:
                    return False
            else:
                pvk = g16.prepare_verifying_key(st.verifying_key)
                snark = sp['snark_proof']
                pub = [st.vmin, st.vmax]
                if pairing_checker is not None:
                    d = g16.calculate_d(pvk, snark, pub)
                    pairing_checker.add_multiple_sources_and_target([snark.a, snark.c, d],
                                                                    [snark.b, pvk.delta_g2_neg_pc, pvk.gamma_g2_neg_pc],
                                                                    pvk.alpha_g1_beta_g2)
                elif not g16.verify_proof(pvk, snark, pub):
                    return False
                resp_for_message = shared.get((sid, 0))
                if resp_for_message is None:
                    return False
                schnorr = SchnorrProtocol(st.schnorr_comm_key(st.verifying_key), snark.d)
                if not schnorr.verify_partial_proof(challenge, sp['sp'], {0: resp_for_message}):
                    return False
        return pairing_checker.verify() if pairing_checker is not None else True
