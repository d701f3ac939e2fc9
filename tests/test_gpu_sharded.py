"""GPU: single-process multi-GPU MSM behind the C ABI (dg_init_devices / dg_msm_*_sharded)
and the hardening of the asynchronous device-pointer entry points.

The reference is ONE process with rayon threads (utils/src/macros.rs:68-84), so the sharded calls are issued from one
host thread and use every GPU the box shows (1 on the single-GPU test box: the same code path with one shard).
sharded == single-GPU == known-discrete-log identity, bit for bit after normalisation.
"""
import numpy as np
import pytest
import torch

from tests import helpers as h

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mdg():
    """The library re-initialised over every visible GPU; restored to cuda:0 afterwards."""
    from crypto_b200 import lib
    ndev = torch.cuda.device_count()
    lib.init(0)
    lib.shutdown()
    lib.init_devices(list(range(ndev)))
    assert lib.device_count() == ndev
    yield lib
    lib.shutdown()
    lib.init(0)


@pytest.mark.parametrize('n', [0, 1, 5, 1000, 1 << 14, (1 << 16) + 3])
def test_sharded_g1_equals_single_and_known_dlog(mdg, cref, n):
    bases, ks = h.g1_bases(max(n, 1), 900 + n)
    ss = h.rand_scalars(max(n, 1), 901 + n)
    bases, ks, ss = bases[:96 * n], ks[:32 * n], ss[:32 * n]
    got = h.affine_g1(mdg.msm_sharded(bases, ss))
    single = h.affine_g1(mdg.msm(bases, ss)) if n else bytes(96)
    assert got == single
    if n:
        assert got == h.known_dlog_msm_g1(ks, ss)


def test_sharded_resident_handle_prefix_and_precompute(mdg, cref):
    n = 1 << 15
    bases, ks = h.g1_bases(n, 77)
    ss = h.rand_scalars(n, 78)
    hb = mdg.ShardedBases(bases)
    exp = h.known_dlog_msm_g1(ks, ss)
    assert h.affine_g1(mdg.msm_sharded(hb, ss)) == exp
    # a prefix shorter than the upload: the trailing shards shrink or go empty
    for m in (n - 1, n // 2 + 7, 3):
        assert h.affine_g1(mdg.msm_sharded(hb, ss[:32 * m])) == h.known_dlog_msm_g1(ks[:32 * m], ss[:32 * m])
    hb.precompute(0)                                    # every device builds the table of its own range
    assert h.affine_g1(mdg.msm_sharded(hb, ss)) == exp
    assert h.affine_g1(mdg.msm_sharded(hb, ss[:32 * 1000])) == h.known_dlog_msm_g1(ks[:32 * 1000], ss[:32 * 1000])
    hb.free()


def test_sharded_g2_and_unchecked(mdg, cref):
    n = 3001
    bases, ks = h.g2_bases(n, 31)
    ss = h.rand_scalars(n, 32)
    assert h.affine_g2(mdg.msm_sharded(bases, ss, g2=True)) == h.known_dlog_msm_g2(ks, ss)
    hb = mdg.ShardedBases(bases, g2=True)
    assert h.affine_g2(mdg.msm_sharded(hb, ss, g2=True)) == h.known_dlog_msm_g2(ks, ss)
    hb.free()


def test_sharded_rejects_non_canonical_scalar_and_recovers(mdg, cref):
    n = 4096
    bases, ks = h.g1_bases(n, 5)
    ss = np.array(h.rand_scalars(n, 6))
    bad = ss.copy()
    bad[32 * (n - 1):32 * n] = 0xff                     # >= r, lands in the last shard
    with pytest.raises(mdg.DockGpuError):
        mdg.msm_sharded(bases, bad)
    assert h.affine_g1(mdg.msm_sharded(bases, ss)) == h.known_dlog_msm_g1(ks, ss)


def test_sharded_calls_from_concurrent_threads(mdg, cref):
    import threading
    n = 5000
    bases, ks = h.g1_bases(n, 11)
    outs, errs = {}, []

    def work(i):
        try:
            ss = h.rand_scalars(n, 100 + i)
            outs[i] = (h.affine_g1(mdg.msm_sharded(bases, ss)), h.known_dlog_msm_g1(ks, ss))
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    assert all(a == b for a, b in outs.values())


# ---- asynchronous device-pointer variants (err_flag, scratch shared between streams) -----------------
def test_device_variant_reports_bad_scalar_through_stream_status(dg, cref):
    n = 2048
    bases, ks = h.g1_bases(n, 41)
    ss = np.array(h.rand_scalars(n, 42))
    bad = ss.copy()
    bad[32 * 7:32 * 8] = 0xff
    dev = torch.device('cuda', 0)
    d_b = torch.from_numpy(np.array(bases)).to(dev)
    d_bad = torch.from_numpy(bad).to(dev)
    d_ok = torch.from_numpy(ss).to(dev)
    d_out = torch.zeros(144, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    dg.msm_device(d_b.data_ptr(), d_bad.data_ptr(), n, d_out.data_ptr(), st.cuda_stream)
    with pytest.raises(dg.DockGpuError):
        dg.stream_status(st.cuda_stream)
    # the flag belongs to that run only: the next valid MSM (async or host path) is clean
    dg.msm_device(d_b.data_ptr(), d_ok.data_ptr(), n, d_out.data_ptr(), st.cuda_stream)
    dg.stream_status(st.cuda_stream)
    assert h.affine_g1(d_out.cpu().numpy()) == h.known_dlog_msm_g1(ks, ss)
    dg.msm_device(d_b.data_ptr(), d_bad.data_ptr(), n, d_out.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert h.affine_g1(dg.msm(bases, ss)) == h.known_dlog_msm_g1(ks, ss)


def test_two_streams_on_one_thread_do_not_corrupt_each_other(dg, cref):
    n = 1 << 16
    dev = torch.device('cuda', 0)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ins = []
    for i in range(2):
        bases, ks = h.g1_bases(n, 300 + i)
        ss = h.rand_scalars(n, 310 + i)
        ins.append((torch.from_numpy(np.array(bases)).to(dev), torch.from_numpy(np.array(ss)).to(dev),
                    torch.zeros(144, dtype=torch.uint8, device=dev), h.known_dlog_msm_g1(ks, ss)))
    torch.cuda.synchronize()
    for rep in range(3):
        for (b, s, o, _), st in zip(ins, (s1, s2)):
            dg.msm_device(b.data_ptr(), s.data_ptr(), n, o.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    for b, s, o, exp in ins:
        assert h.affine_g1(o.cpu().numpy()) == exp
