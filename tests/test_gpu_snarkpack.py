"""GPU: SnarkPack aggregation of Groth16 proofs (row f3) re-expressing legogroth16/src/aggregation/tests.rs
(groth16_aggregation): n proofs of one circuit -> aggregate_proofs -> verify_aggregate_proof; a wrong public input, a
proof that is not in the aggregate, a tampered aggregate and a wrong transcript are rejected.  Every pairing product,
MSM, key folding and fixed-base multiplication behind it runs on the GPU through the C ABI; the pieces that have a closed
form (SRS powers, key compression, pair commitments) are also checked against the oracle."""
import random

import numpy as np
import pytest

from crypto_b200 import groth16 as g16
from crypto_b200 import group as gp
from crypto_b200 import proof_system as ps
from crypto_b200 import snarkpack as sp
from oracle import bls12_381 as o
from tests import helpers as h

pytestmark = pytest.mark.gpu

R = o.R


def _lego_proofs(n, rng):
    """n LegoGroth16 proofs of the bound-check circuit with ONE committed witness (the value) and v != 0, so D is a real
    commitment (legogroth16/src/aggregation/tests.rs: legogroth16_aggregation)."""
    cs = ps.bound_check_circuit(16)
    pk, ni = g16.generate_parameters(cs, *(rng.randrange(1, R) for _ in range(5)), t=rng.randrange(1 << 100, 1 << 200),
                                     g1_generator=o.g1_to_bytes(o.G1_GEN), g2_generator=o.g2_to_bytes(o.G2_GEN), commit_witness_count=1)
    dpk = g16.DeviceProvingKey(pk, cs)
    pvk = g16.prepare_verifying_key(pk.vk)
    proofs, inputs = [], []
    for i in range(n):
        vmin, vmax = 20 + i, 9000 + 11 * i
        value = rng.randrange(vmin, vmax)
        proof, _ = g16.create_proof(dpk, ps.bound_check_assignment(value, vmin, vmax, 16), rng.randrange(1, R), rng.randrange(1, R),
                                    rng.randrange(1, R))
        assert g16.verify_proof(pvk, proof, [vmin, vmax])
        assert not gp.is_identity(proof.d)
        proofs.append((proof.a, proof.b, proof.c, proof.d))
        inputs.append([vmin, vmax])
    dpk.free()
    g16.KEY_CACHE.clear()
    return pk.vk, proofs, inputs


def _proofs(n, rng):
    cs = ps.bound_check_circuit(16)
    pk, ni = g16.generate_parameters(cs, *(rng.randrange(1, R) for _ in range(5)), t=rng.randrange(1 << 100, 1 << 200),
                                     g1_generator=o.g1_to_bytes(o.G1_GEN), g2_generator=o.g2_to_bytes(o.G2_GEN), commit_witness_count=0)
    dpk = g16.DeviceProvingKey(pk, cs)
    pvk = g16.prepare_verifying_key(pk.vk)
    proofs, inputs = [], []
    for i in range(n):
        vmin, vmax = 10 + i, 5000 + 7 * i
        value = rng.randrange(vmin, vmax)
        proof, _ = g16.create_proof(dpk, ps.bound_check_assignment(value, vmin, vmax, 16), rng.randrange(1, R), rng.randrange(1, R), 0)
        assert g16.verify_proof(pvk, proof, [vmin, vmax])           # v = 0, no committed witnesses: a plain Groth16 proof
        assert gp.is_identity(proof.d)
        proofs.append((proof.a, proof.b, proof.c))
        inputs.append([vmin, vmax])
    dpk.free()
    g16.KEY_CACHE.clear()
    return pk.vk, proofs, inputs


def test_srs_keys_and_commitments_against_the_oracle(dg, cref):
    rng = random.Random(3)
    alpha, beta, n = rng.randrange(1, R), rng.randrange(1, R), 4
    g, hh = o.g1_to_bytes(o.G1_GEN), o.g2_to_bytes(o.G2_GEN)
    srs = sp.setup_fake_srs(alpha, beta, n, g, hh)
    assert b''.join(srs.g_alpha_powers) == bytes(cref.g1_generator_muls(h.scalars_bytes(sp.powers(alpha, 2 * n))))
    assert b''.join(srs.h_beta_powers) == bytes(cref.g2_generator_muls(h.scalars_bytes(sp.powers(beta, 2 * n))))
    psrs, vsrs = srs.specialize(n)
    assert psrs.has_correct_len(n) and vsrs.g == g and vsrs.h_alpha == srs.h_alpha_powers[1]
    # Key::compress / scale by closed form: w_i = g^(alpha^(n+i))
    x = rng.randrange(1, R)
    left, right = psrs.wkey.split(n // 2)
    comp = left.compress(right, x)
    exp = [(pow(alpha, n + i, R) + x * pow(alpha, n + n // 2 + i, R)) % R for i in range(n // 2)]
    assert b''.join(comp.a) == bytes(cref.g1_generator_muls(h.scalars_bytes(exp)))
    s_vec = [rng.randrange(1, R) for _ in range(n)]
    scaled = psrs.vkey.scale(s_vec)
    assert b''.join(scaled.b) == bytes(cref.g2_generator_muls(h.scalars_bytes([pow(beta, i, R) * s_vec[i] % R for i in range(n)])))
    # PairCommitment::double against the oracle's multi_pairing
    a = [bytes(x) for x in sp._split_records(cref.g1_generator_muls(h.rand_scalars(n, 5)), 96)]
    b = [bytes(x) for x in sp._split_records(cref.g2_generator_muls(h.rand_scalars(n, 6)), 192)]
    t, u = sp._pairing_products(sp.commit_double_products(psrs.vkey, psrs.wkey, a, b))
    cat = lambda v: np.frombuffer(b''.join(v), dtype=np.uint8)
    assert t == bytes(cref.multi_pairing(cat(a + psrs.wkey.a), cat(psrs.vkey.a + b)))
    assert u == bytes(cref.multi_pairing(cat(a + psrs.wkey.b), cat(psrs.vkey.b + b)))
    # utils::compress in G2
    out = sp.compress(b, n // 2, x, gp.G2)
    e0 = o.E2.add(o.g2_from_bytes(b[0]), o.E2.mul(o.g2_from_bytes(b[n // 2]), x))
    assert out[0] == o.g2_to_bytes(e0)


@pytest.mark.parametrize('n', [2, 8])
def test_groth16_aggregation(dg, cref, n):
    rng = random.Random(100 + n)
    vk, proofs, inputs = _proofs(n, rng)
    srs = sp.setup_fake_srs(rng.randrange(1, R), rng.randrange(1, R), n, o.g1_to_bytes(o.G1_GEN), o.g2_to_bytes(o.G2_GEN))
    psrs, vsrs = srs.specialize(n)
    agg = sp.aggregate_proofs(psrs, sp.Transcript(b'test'), proofs)
    assert agg.gipa.nproofs == n and len(agg.gipa.comms_ab) == n.bit_length() - 1
    # z_c = sum r^i C_i and z_ab = prod e(A_i, B_i)^(r^i): recompute r from the transcript and check by the oracle
    tr = sp.Transcript(b'test')
    tr.append(b'AB-commitment', agg.com_ab); tr.append(b'C-commitment', agg.com_c)
    r = tr.challenge_scalar(b'r-random-fiatshamir')
    acc = None
    for i, (_, _, c) in enumerate(proofs):
        acc = o.E1.add(acc, o.E1.mul(o.g1_from_bytes(c), pow(r, i, R)))
    assert agg.z_c == o.g1_to_bytes(acc)
    assert sp.verify_aggregate_proof(vsrs, vk, inputs, agg, sp.Transcript(b'test'), 0xfeedface12345)
    assert sp.verify_aggregate_proof(vsrs, vk, inputs, agg, sp.Transcript(b'test'), 0xfeedface12345, lazy=False)
    # rejections
    bad_inputs = [list(x) for x in inputs]
    bad_inputs[n - 1][0] += 1
    assert not sp.verify_aggregate_proof(vsrs, vk, bad_inputs, agg, sp.Transcript(b'test'), 7)
    assert not sp.verify_aggregate_proof(vsrs, vk, inputs, agg, sp.Transcript(b'other'), 7)
    swapped = list(proofs)
    swapped[0], swapped[1] = swapped[1], swapped[0]
    agg_sw = sp.aggregate_proofs(psrs, sp.Transcript(b'test'), swapped)
    assert not sp.verify_aggregate_proof(vsrs, vk, inputs, agg_sw, sp.Transcript(b'test'), 7)
    tampered = sp.AggregateProof(agg.com_ab, agg.com_c, agg.z_ab, gp.mul_affine(agg.z_c, 2), agg.gipa, agg.vkey_opening, agg.wkey_opening)
    assert not sp.verify_aggregate_proof(vsrs, vk, inputs, tampered, sp.Transcript(b'test'), 7)
    with pytest.raises(ValueError):
        sp.aggregate_proofs(psrs, sp.Transcript(b'test'), proofs[:-1] if n > 2 else proofs[:1])


@pytest.mark.parametrize('n', [2, 4])
def test_legogroth16_aggregation(dg, cref, n):
    """aggregation/legogroth16: A, B, C AND D committed and folded (AggregateLegoProof); z_d = sum r^i D_i by the oracle; a
    wrong input, a wrong transcript, a tampered z_d and a Groth16-shaped aggregate are rejected."""
    rng = random.Random(300 + n)
    vk, proofs, inputs = _lego_proofs(n, rng)
    srs = sp.setup_fake_srs(rng.randrange(1, R), rng.randrange(1, R), n, o.g1_to_bytes(o.G1_GEN), o.g2_to_bytes(o.G2_GEN))
    psrs, vsrs = srs.specialize(n)
    agg = sp.aggregate_lego_proofs(psrs, sp.Transcript(b'lego'), proofs)
    assert agg.is_lego and len(agg.gipa.comms_d) == n.bit_length() - 1 and len(agg.gipa.z_d) == n.bit_length() - 1
    tr = sp.Transcript(b'lego')
    tr.append(b'AB-commitment', agg.com_ab); tr.append(b'C-commitment', agg.com_c); tr.append(b'D-commitment', agg.com_d)
    r = tr.challenge_scalar(b'r-random-fiatshamir')
    acc = None
    for i, p in enumerate(proofs):
        acc = o.E1.add(acc, o.E1.mul(o.g1_from_bytes(p[3]), pow(r, i, R)))
    assert agg.z_d == o.g1_to_bytes(acc)
    # com_d = (prod e(D_i, v1_i), prod e(D_i, v2_i)) by the oracle's multi_pairing
    cat = lambda v: np.frombuffer(b''.join(v), dtype=np.uint8)
    d = [p[3] for p in proofs]
    assert agg.com_d[0] == bytes(cref.multi_pairing(cat(d), cat(psrs.vkey.a)))
    assert agg.com_d[1] == bytes(cref.multi_pairing(cat(d), cat(psrs.vkey.b)))
    assert sp.verify_aggregate_lego_proof(vsrs, vk, inputs, agg, sp.Transcript(b'lego'), 0xabcdef0123)
    assert sp.verify_aggregate_lego_proof(vsrs, vk, inputs, agg, sp.Transcript(b'lego'), 0xabcdef0123, lazy=False)
    bad_inputs = [list(x) for x in inputs]
    bad_inputs[0][1] += 1
    assert not sp.verify_aggregate_lego_proof(vsrs, vk, bad_inputs, agg, sp.Transcript(b'lego'), 9)
    assert not sp.verify_aggregate_lego_proof(vsrs, vk, inputs, agg, sp.Transcript(b'other'), 9)
    import dataclasses
    assert not sp.verify_aggregate_lego_proof(vsrs, vk, inputs, dataclasses.replace(agg, z_d=gp.mul_affine(agg.z_d, 3)), sp.Transcript(b'lego'), 9)
    bad_gipa = dataclasses.replace(agg.gipa, final_d=gp.mul_affine(agg.gipa.final_d, 2))
    assert not sp.verify_aggregate_lego_proof(vsrs, vk, inputs, dataclasses.replace(agg, gipa=bad_gipa), sp.Transcript(b'lego'), 9)
    # the Groth16 verifier on a LegoGroth16 aggregate of proofs with D != 0 has no term for D and must not accept it either way
    g16_agg = sp.aggregate_proofs(psrs, sp.Transcript(b'lego'), [p[:3] for p in proofs])
    assert not sp.verify_aggregate_lego_proof(vsrs, vk, inputs, g16_agg, sp.Transcript(b'lego'), 9)
    assert not sp.verify_aggregate_proof(vsrs, vk, inputs, g16_agg, sp.Transcript(b'lego'), 9)


def test_legogroth16_aggregation_using_groth16(dg, n=4):
    """aggregation/legogroth16/using_groth16.rs: the Groth16 aggregate of (A, B, C) plus the D's in the clear."""
    rng = random.Random(404)
    vk, proofs, inputs = _lego_proofs(n, rng)
    srs = sp.setup_fake_srs(rng.randrange(1, R), rng.randrange(1, R), n, o.g1_to_bytes(o.G1_GEN), o.g2_to_bytes(o.G2_GEN))
    psrs, vsrs = srs.specialize(n)
    agg, d = sp.aggregate_lego_proofs_using_groth16(psrs, sp.Transcript(b'u'), proofs)
    assert not agg.is_lego and d == [p[3] for p in proofs]
    assert sp.verify_aggregate_lego_proof_using_groth16(vsrs, vk, inputs, agg, d, sp.Transcript(b'u'), 0x777)
    d_bad = list(d)
    d_bad[1] = gp.mul_affine(d_bad[1], 2)
    assert not sp.verify_aggregate_lego_proof_using_groth16(vsrs, vk, inputs, agg, d_bad, sp.Transcript(b'u'), 0x777)
    assert not sp.verify_aggregate_lego_proof_using_groth16(vsrs, vk, inputs, agg, d[:-1], sp.Transcript(b'u'), 0x777)
    bad_inputs = [list(x) for x in inputs]
    bad_inputs[2][0] -= 1
    assert not sp.verify_aggregate_lego_proof_using_groth16(vsrs, vk, bad_inputs, agg, d, sp.Transcript(b'u'), 0x777)
