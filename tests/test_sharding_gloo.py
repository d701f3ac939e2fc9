"""CPU, world_size 2 over gloo: the multi-GPU host logic of SURVEY.md 8e / bench.py --
contiguous base-range shards, ONE all-gather of the 144-byte partial results, fold under the
group law -- with the oracle standing in for the per-rank GPU MSM (no GPU in this container)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crypto_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import cref
    ks = cref.random_scalars(n, 5)
    ss = cref.random_scalars(n, 6)
    bases = cref.g1_generator_muls(ks)
    lo, hi = sharding.shard_range(n, rank, world)
    part = cref.msm_g1(bases[96 * lo:96 * hi], ss[32 * lo:32 * hi], hi - lo)

    def fold(parts):                       # oracle stand-in for dg_fold_g1
        acc = None
        from oracle import bls12_381 as o
        for i in range(len(parts) // 144):
            aff = bytes(cref.normalize_batch_g1(parts[144 * i:144 * i + 144]))
            acc = o.E1.add(acc, o.g1_from_bytes(aff))
        return o.g1_to_bytes(acc)

    total = sharding.all_gather_fold(torch.from_numpy(np.array(part)), fold)
    if rank == 0:
        full = bytes(cref.normalize_batch_g1(cref.msm_g1(bases, ss)))
        q.put(total == full)
    dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 3):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_msm_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 300, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
