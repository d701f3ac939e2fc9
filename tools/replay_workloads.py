"""Replays the L0 call sequences of BASELINE.json configs 3 and 4 through the
C ABI with synthetic inputs of the same shapes, checks every result (known-dlog identity / oracle)
and times the GPU next to the CPU oracle.  Writes one JSON object (stdout or --out FILE).

  config 3  legogroth16::create_random_proof, D = 2^18: 4 G1 MSMs of ~2^18 terms over the
            proving-key queries + 1 G2 MSM (b_g2_query) + a small gamma_abc MSM
            (legogroth16/src/prover.rs:286,299,326,333,344,361)
  config 4  bbs_plus SignatureG1::new / PoK with 10 000 messages (MSMs of 10 001 terms,
            bbs_plus/src/setup.rs:145, proof.rs:187,241), verification 2-pair product check
            (proof.rs:494), and a vb_accumulator batch witness update for 10 000 members
            (witness.rs:269-284: per-witness mul_bigint + WindowTable multiply + normalize_batch)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cref  # noqa: E402  (checker + CPU timing only)
from crypto_b200 import lib, msm  # noqa: E402

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def ints(b):
    return [int.from_bytes(bytes(r), 'little') for r in np.asarray(b, dtype=np.uint8).reshape(-1, 32)]


def sbytes(v):
    return np.frombuffer(b''.join(int(x % R).to_bytes(32, 'little') for x in v), dtype=np.uint8)


def dlog_total(ks, ss):
    return sum(a * b for a, b in zip(ints(ks), ints(ss))) % R


def timeit(fn, reps=3):
    fn()
    best = 1e9
    out = None
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return best, out


def config3(logd, with_cpu):
    n = 1 << logd
    res = {'D': n, 'msms': []}
    total_gpu = total_cpu = 0.0
    for name, seed, g2 in (('h_query x h', 1, False), ('l_query x witness', 2, False), ('a_query', 3, False),
                           ('b_g1_query', 4, False), ('b_g2_query', 5, True)):
        ks = cref.random_scalars(n, 1000 + seed)
        ss = cref.random_scalars(n, 2000 + seed).copy()
        sv = ss.reshape(n, 32)
        # real witnesses are full of 0/1 and small values (SURVEY 7 'hard parts'): make 30% of them so
        sv[: n // 10] = 0
        sv[n // 10: n // 5] = 0
        sv[n // 10: n // 5, 0] = 1
        sv[n // 5: 3 * n // 10, 2:] = 0
        bases = cref.g2_generator_muls(ks) if g2 else cref.g1_generator_muls(ks)
        hb = lib.Bases(bases, g2=g2).precompute()
        dt, out = timeit(lambda: lib.msm(hb, ss, g2=g2))
        hb.free()
        tot = dlog_total(ks, ss)
        if g2:
            ok = bytes(cref.normalize_batch_g2(out)) == bytes(cref.g2_generator_muls(sbytes([tot])))
        else:
            ok = bytes(cref.normalize_batch_g1(out)) == bytes(cref.g1_generator_muls(sbytes([tot])))
        ent = {'name': name, 'group': 'G2' if g2 else 'G1', 'terms': n, 'gpu_ms': dt * 1e3, 'ok': ok}
        total_gpu += dt
        if with_cpu:
            t = time.perf_counter()
            (cref.msm_g2 if g2 else cref.msm_g1)(bases, ss)
            ent['cpu_ms'] = (time.perf_counter() - t) * 1e3
            total_cpu += ent['cpu_ms'] / 1e3
        res['msms'].append(ent)
        print('  ', ent, flush=True)
    res['gpu_ms_total_msm'] = total_gpu * 1e3
    if with_cpu:
        res['cpu_ms_total_msm'] = total_cpu * 1e3
    return res


def config4(nmsg, with_cpu):
    res = {'messages': nmsg}
    hs, hk = None, None
    ks = cref.random_scalars(nmsg + 1, 31)
    hs = cref.g1_generator_muls(ks)
    ss = cref.random_scalars(nmsg + 1, 32)
    hb = lib.Bases(hs)
    dt, out = timeit(lambda: lib.msm(hb, ss))
    ok = bytes(cref.normalize_batch_g1(out)) == bytes(cref.g1_generator_muls(sbytes([dlog_total(ks, ss)])))
    res['sign_msm'] = {'terms': nmsg + 1, 'gpu_ms': dt * 1e3, 'ok': ok}
    if with_cpu:
        t = time.perf_counter(); cref.msm_g1(hs, ss); res['sign_msm']['cpu_ms'] = (time.perf_counter() - t) * 1e3
    hb.free()
    # verification-shaped 2-pair product check  e(A, pk + e g2) * e(-b, g2) == 1
    b_aff = bytes(cref.normalize_batch_g1(out))
    e, x = 0x2222, 0x3333
    A = bytes(cref.normalize_batch_g1(cref.batch_mul_g1(np.frombuffer(b_aff, np.uint8), sbytes([pow(e + x, -1, R)]))))
    nb = bytes(cref.normalize_batch_g1(cref.batch_mul_g1(np.frombuffer(b_aff, np.uint8), sbytes([R - 1]))))
    g2 = bytes(cref.g2_generator_muls(sbytes([1])))
    pk = bytes(cref.g2_generator_muls(sbytes([x + e])))
    dt, r = timeit(lambda: lib.multi_pairing_is_one(A + nb, pk + g2))
    res['verify_2pair_check'] = {'gpu_ms': dt * 1e3, 'ok': bool(r)}
    if with_cpu:
        t = time.perf_counter(); cref.multi_pairing(A + nb, pk + g2); res['verify_2pair_check']['cpu_ms'] = (time.perf_counter() - t) * 1e3
    # many pairs as the lazy RandomizedPairingChecker produces them
    k = 256
    ps = cref.g1_generator_muls(cref.random_scalars(k, 41)); qs = cref.g2_generator_muls(cref.random_scalars(k, 42))
    dt, ml = timeit(lambda: lib.multi_pairing(ps, qs))
    ok = bytes(ml) == bytes(cref.multi_pairing(ps, qs))
    res['multi_pairing_256'] = {'pairs': k, 'gpu_ms': dt * 1e3, 'ok': ok}
    if with_cpu:
        t = time.perf_counter(); cref.multi_pairing(ps, qs); res['multi_pairing_256']['cpu_ms'] = (time.perf_counter() - t) * 1e3
    # accumulator batch witness update
    m = nmsg
    wits = cref.g1_generator_muls(cref.random_scalars(m, 51))
    v = cref.g1_generator_muls(cref.random_scalars(1, 52))
    sa, sb = cref.random_scalars(m, 53), cref.random_scalars(m, 54)

    def gpu_update():
        t = lib.FixedBaseTable(v, m)
        o = lib.batch_mul_add_fixed_g1(wits, sa, t, sb)
        t.free()
        return o
    dt, out = timeit(gpu_update)
    res['witness_update'] = {'witnesses': m, 'gpu_ms': dt * 1e3}
    t = time.perf_counter()
    left = cref.batch_mul_g1(wits, sa)
    right, _, _ = cref.fixed_base_mul_many_g1(v, m, sb)
    cpu_s = time.perf_counter() - t
    # check the first 16 against the oracle sum
    from oracle import bls12_381 as o
    la = bytes(cref.normalize_batch_g1(left[:144 * 16])); ra = bytes(cref.normalize_batch_g1(right[:144 * 16]))
    exp = b''.join(o.g1_to_bytes(o.E1.add(o.g1_from_bytes(la[96 * i:96 * i + 96]), o.g1_from_bytes(ra[96 * i:96 * i + 96]))) for i in range(16))
    res['witness_update']['ok'] = bytes(out[:96 * 16]) == exp
    res['witness_update']['cpu_ms'] = cpu_s * 1e3
    return res


def config_generator(logd, with_cpu):
    """'Next' row f2: legogroth16 CRS generation (legogroth16/src/generator.rs:335-425): FixedBase::msm of
    one generator over D scalars, five times in G1 and once in G2, each followed by normalize_batch."""
    n = 1 << logd
    res = {'D': n, 'tables': []}
    for name, g2, seed in (('a_query', False, 1), ('b_g1_query', False, 2), ('h_query', False, 3), ('l_query', False, 4),
                           ('gamma_abc', False, 5), ('b_g2_query', True, 6)):
        base = (cref.g2_generator_muls if g2 else cref.g1_generator_muls)(cref.random_scalars(1, 60 + seed))
        ss = cref.random_scalars(n, 70 + seed).copy()
        ss[:64] = 0                                   # scalar 0 -> identity points in the queries (generator.rs:342,373)
        t = lib.FixedBaseTable(base, n, g2=g2)
        dt, out = timeit(lambda: t.mul_many_normalized(ss), reps=2)
        t.free()
        k = 64                                        # check the first 64 against per-scalar mul_bigint on the oracle
        exp = (cref.normalize_batch_g2(cref.batch_mul_g2(np.tile(base, k), ss[:32 * k])) if g2
               else cref.normalize_batch_g1(cref.batch_mul_g1(np.tile(base, k), ss[:32 * k])))
        rec = 192 if g2 else 96
        ent = {'name': name, 'group': 'G2' if g2 else 'G1', 'scalars': n, 'gpu_ms': dt * 1e3, 'ok': bytes(out[:rec * k]) == bytes(exp)}
        if with_cpu:
            tt = time.perf_counter()
            if g2:
                j, _, _ = cref.fixed_base_mul_many_g2(base, n, ss); cref.normalize_batch_g2(j)
            else:
                j, _, _ = cref.fixed_base_mul_many_g1(base, n, ss); cref.normalize_batch_g1(j)
            ent['cpu_ms'] = (time.perf_counter() - tt) * 1e3
        res['tables'].append(ent)
        print('  ', ent, flush=True)
    res['gpu_ms_total'] = sum(e['gpu_ms'] for e in res['tables'])
    if with_cpu:
        res['cpu_ms_total'] = sum(e['cpu_ms'] for e in res['tables'])
    return res


def config_snarkpack(nproofs, with_cpu):
    """'Next' row f3: SnarkPack TIPP/MIPP prover shape (legogroth16/src/aggregation/utils.rs:85-97,
    commitment.rs:40-66, groth16/prover.rs:114-116): per GIPA round with half-size m, four multi_pairings of
    2m pairs (two PairCommitment::double) and two of m pairs, then one MSM of n terms."""
    n = nproofs
    a = cref.g1_generator_muls(cref.random_scalars(2 * n, 81))
    b = cref.g2_generator_muls(cref.random_scalars(2 * n, 82))
    res = {'proofs': n, 'rounds': []}
    gpu_total = cpu_total = 0.0
    pairs_total = 0
    m = n // 2
    while m >= 1:
        sizes = [2 * m, 2 * m, 2 * m, 2 * m, m, m]
        t = time.perf_counter()
        outs = lib.multi_pairing_batch(np.concatenate([a[:96 * k] for k in sizes]), np.concatenate([b[:192 * k] for k in sizes]), sizes)
        g = time.perf_counter() - t
        ent = {'m': m, 'pairs': sum(sizes), 'gpu_ms': g * 1e3}
        gpu_total += g
        pairs_total += sum(sizes)
        if with_cpu:
            t = time.perf_counter()
            exp = [cref.multi_pairing(a[:96 * k], b[:192 * k]) for k in sizes]
            c = time.perf_counter() - t
            ent['cpu_ms'] = c * 1e3
            cpu_total += c
            ent['ok'] = all(bytes(x) == bytes(y) for x, y in zip(outs, exp))
        res['rounds'].append(ent)
        m //= 2
    ks = cref.random_scalars(n, 83); ss = cref.random_scalars(n, 84)
    cpts = cref.g1_generator_muls(ks)
    dt, out = timeit(lambda: lib.msm(cpts, ss))
    res['z_c_msm'] = {'terms': n, 'gpu_ms': dt * 1e3,
                      'ok': bytes(cref.normalize_batch_g1(out)) == bytes(cref.g1_generator_muls(sbytes([dlog_total(ks, ss)])))}
    res['pairs_total'] = pairs_total
    res['gpu_ms_total_pairings'] = gpu_total * 1e3
    if with_cpu:
        res['cpu_ms_total_pairings'] = cpu_total * 1e3
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--logd', type=int, default=18)
    ap.add_argument('--messages', type=int, default=10000)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    lib.init()
    out = {'host_cores': len(os.sched_getaffinity(0)), 'note': 'gpu_ms = host C-ABI call incl. H2D of scalars and D2H of the result; '
           'cpu_ms = oracle C restatement of the arkworks algorithm on the host cores'}
    print('config 3', flush=True)
    out['config3_legogroth16_prover_shape'] = config3(a.logd, not a.no_cpu)
    print('config 4', flush=True)
    out['config4_bbs_plus_and_accumulator_shape'] = config4(a.messages, not a.no_cpu)
    print('generator (row f2)', flush=True)
    out['next_f2_crs_generator_shape'] = config_generator(a.logd, not a.no_cpu)
    print('snarkpack (row f3)', flush=True)
    out['next_f3_snarkpack_prover_shape'] = config_snarkpack(256, not a.no_cpu)
    s = json.dumps(out, indent=1)
    if a.out:
        open(a.out, 'w').write(s)
    print(s)


if __name__ == '__main__':
    main()
