#!/usr/bin/env python3
"""bench.py -- BLS12-381 G1 MSM throughput (BASELINE.json metric) on N x B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--logn 20] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...          (one rank per GPU, NCCL)

A "step" is one G1 MSM over synthetic seeded scalars/bases (bases k_i*G with known k_i, scalars
uniform in [0, r)); every timed configuration first passes the known-discrete-log identity
sum s_i (k_i G) = (sum s_i k_i mod r) G, bit-exact in affine form.

`value`   : terms/s with scalars and the RAW 96-byte bases already resident in HBM
            (dg_msm_g1_device): like for like with msm_bigint(bases, scalars), nothing precomputed.
            N = 1: one 2^20-term MSM (the size the metric is quoted on).  N > 1: weak scaling, every
            rank owns a contiguous 2^20-term base range of one N*2^20-term MSM; the 144-byte partial
            results are all-gathered (NCCL) and folded on the GPU under the group law.
`value_resident_table`: the same MSM through a handle whose bases carry the 2^(ck)-multiples table
            built once at upload (dg_bases_precompute; the proving-key mode).
`e2e`     : terms/s through the host C-ABI call with the scalars in pinned HOST memory copied every
            step and the 144-byte result read back every step (bases resident behind a plain handle).
            N = 1: dg_msm_g1.  N > 1: ONE call of dg_msm_g1_sharded from ONE host thread of rank 0
            driving all N GPUs (the reference is one process), the other ranks idle.
`sweep`   : (N = 1) 2^16 .. 2^24 terms on one GPU, raw bases and resident table (BASELINE config 2).
`strong_2p24`: one 2^24-term MSM (BASELINE config 5): N = 1 on one GPU; N > 1 split by base range over
            the ranks (device-resident, NCCL all-gather + fold), next to the same MSM on rank 0's
            GPU alone measured in the same run, and through dg_msm_g1_sharded from host memory.
`roofline`: the dominant kernel of the `value` path (round 0 of the batch-affine stage) against the
            measured HBM peak, algorithmic bytes 128 B/term; `int_roofline` /
            `mult_roofline` are the integer-pipe readings of the same launch.
`cpu_baseline`: the oracle's C restatement of the arkworks rayon algorithm on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'BLS12-381 G1 MSM scalar-muls/s'
UNIT = 'scalar-muls/s'
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SWEEP = (16, 18, 20, 22, 24)


def bench_config(logn, world):
    """The `config` both arms print (identical dicts: the driver compares them)."""
    n = 1 << logn
    return {'workload': 'bls12-381 g1 msm, 2^%d random scalars/bases per GPU' % logn, 'terms_per_gpu': n,
            'global_terms': n * world, 'n_gpus': world}


def seed_of(n, rank=0):
    return (0xD0C4C0DE ^ n) + 7919 * rank


def synth_scalars(n, rank=0):
    """Seeded inputs: scalars uniform in [0, r) and the discrete logs k_i of the bases
    (oracle helper = input synthesis, not the measured path)."""
    from oracle import cref
    seed = seed_of(n, rank)
    return cref.random_scalars(n, seed), cref.random_scalars(n, seed + 1)


def cpu_bases(ks):
    from oracle import cref
    return cref.g1_generator_muls(ks)


def gpu_bases(lib, ks):
    """k_i * G on the GPU through the library's own fixed-base path (verified against the oracle in
    tests/): input synthesis for sizes where the CPU helper would take minutes."""
    from oracle import cref
    one = np.zeros(32, np.uint8)
    one[0] = 1
    g = cref.g1_generator_muls(one)
    n = len(ks) // 32
    tbl = lib.FixedBaseTable(g, max(n, 32))
    out = np.array(tbl.mul_many_normalized(ks))
    tbl.free()
    return out


def expected_point(ks, ss):
    """(sum s_i k_i mod r) * G as affine bytes."""
    from oracle import cref
    tot = cref.scalar_dot_mod_r(ks, ss)
    return bytes(cref.g1_generator_muls(np.frombuffer(tot.to_bytes(32, 'little'), dtype=np.uint8)))


def affine_of(jac):
    from oracle import cref
    return bytes(cref.normalize_batch_g1(np.asarray(jac, dtype=np.uint8)))


class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device, period_ms=200):
        self.device = device
        self.proc = None
        self.period_ms = period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        out = self.proc.communicate()[0]
        sm, smax, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append((float(f[1]), float(f[3]))); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower() == 'active':
                    reasons.add(name)
        # the sampler also sees the idle gaps between timed regions (host-side input synthesis): "under load" = samples
        # drawing at least 60 % of the highest power seen
        pmax = max((p for _, p in sm), default=0.0)
        load = [c for c, p in sm if p >= 0.6 * pmax]
        return {'sm_mhz': float(np.median(load)) if load else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm), 'samples_under_load': len(load), 'power_w_max': pmax}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs must use every host core.  Called before the
    oracle library (libgomp) is loaded."""
    cores = host_cores()
    os.environ['OMP_NUM_THREADS'] = str(cores)
    return cores


def time_cpu_msm(bases, scalars, n, reps):
    from oracle import cref
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        out = cref.msm_g1(bases, scalars, n)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best, out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle C restatement of arkworks
    msm_bigint_wnaf, OpenMP over windows like rayon) on the host cores, same config/metric.
    N > 1: rank 0 alone; a step stays one 2^logn-term MSM (a 1/N sample of the N*2^logn workload)."""
    if rank != 0:
        return
    cores = use_all_host_cores()
    from oracle import cref
    cref.set_threads(cores)
    n = 1 << args.logn
    ss, ks = synth_scalars(n)
    bases = cpu_bases(ks)
    for _ in range(min(args.warmup, 1)):
        time_cpu_msm(bases, ss, n, 1)
    times = []
    for _ in range(args.steps):
        dt, _ = time_cpu_msm(bases, ss, n, 1)
        times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': bench_config(args.logn, world),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d x one full 2^%d-term MSM%s (C restatement of ark-ec 0.4 msm_bigint_wnaf, OpenMP over '
                                   'windows; the Rust reference cannot be built in this image)'
                                   % (len(times), args.logn, '' if world == 1 else ' = 1/%d of the %d-GPU workload' % (world, world))},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def load_json(path):
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except ValueError:
            return None
    return None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from crypto_b200 import lib

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    # Rank 0 also drives every GPU of the box through the single-process sharded entry points (its primary device
    # stays its own GPU); the other ranks own one device each.
    single_process = world > 1 and rank == 0 and torch.cuda.device_count() >= world and not args.no_single_process
    if single_process:
        lib.init_devices([local_rank] + [d for d in range(world) if d != local_rank])
    else:
        lib.init(local_rank)
    cpu_group = dist.new_group(backend='gloo') if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    d_out = torch.zeros(144, dtype=torch.uint8, device=dev)
    d_gather = torch.zeros(144 * world, dtype=torch.uint8, device=dev)
    d_final = torch.zeros(144, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        """Host-side barrier (gloo): ranks waiting here leave their GPUs idle."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def timed(step, steps, warmup, collective=True):
        """Total device time (ms) of `steps` calls of step(): CUDA events on the launching stream, L2 flushed
        before every timed call (outside the event pair), max over ranks."""
        for _ in range(warmup):
            flush.fill_(1)
            step()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier() if collective else torch.cuda.synchronize()
        for e0, e1 in evs:
            flush.fill_(1)
            e0.record(stream)
            step()
            e1.record(stream)
        barrier() if collective else torch.cuda.synchronize()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        if world > 1 and collective:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def msm_step(n, d_scalars, d_bases=None, handle=None, combine=True):
        def step():
            if handle is not None:
                lib.msm_handle_device(handle, d_scalars.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)
            else:
                lib.msm_device(d_bases.data_ptr(), d_scalars.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)
            if world > 1 and combine:
                dist.all_gather_into_tensor(d_gather, d_out)
                lib.fold_g1_device(d_gather.data_ptr(), world, d_final.data_ptr(), stream.cuda_stream)
        return step

    def check(step, expected, what, combined=False):
        flush.fill_(1)
        step()
        torch.cuda.synchronize()
        lib.stream_status(stream.cuda_stream)
        got = affine_of((d_final if combined else d_out).cpu().numpy())
        if got != expected:
            raise SystemExit('bench: GPU MSM result differs from the known-dlog identity (%s)' % what)

    def time_host_calls(call, steps):
        """Wall clock around `steps` host C-ABI calls (each returns after its result is back in host memory)."""
        for _ in range(max(5, args.warmup)):          # also brings the PCIe link and the pinned-copy path out of idle
            call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = call()
        return time.perf_counter() - t0, out

    sampler = ClockSampler(local_rank, args.clock_sample_ms)

    # ---------------------------------------------------------------- part A: the headline (2^logn terms per GPU)
    n = 1 << args.logn
    ss, ks = synth_scalars(n, rank)
    bases = gpu_bases(lib, ks)
    expected = expected_point(ks, ss)
    d_bases = torch.from_numpy(bases).to(dev)
    d_scalars = torch.from_numpy(np.array(ss)).to(dev)
    hb = lib.Bases(bases)
    hb_pre = lib.Bases(bases).precompute(args.precompute_window)
    step_plain = msm_step(n, d_scalars, d_bases=d_bases)
    step_pre = msm_step(n, d_scalars, handle=hb_pre)
    check(msm_step(n, d_scalars, d_bases=d_bases, combine=False), expected, 'raw bases')
    check(msm_step(n, d_scalars, handle=hb_pre, combine=False), expected, 'resident table')
    if rank == 0:
        sampler.start()
        time.sleep(0.3)                      # let nvidia-smi come up before the first timed region
    pre_ms = timed(step_pre, args.steps, args.warmup)
    lib.prof_enable(True)
    lib.prof_read_accumulate()
    for _ in range(args.warmup):
        flush.fill_(1)
        step_plain()
    barrier()
    lib.prof_read_accumulate()
    launches0 = lib.launch_count()
    plain_ms = timed(step_plain, args.steps, 0)
    launches = lib.launch_count() - launches0
    dom_ms, dom_cnt = lib.prof_read_accumulate()
    lib.prof_enable(False)
    value = n * world * args.steps / (plain_ms * 1e-3)
    value_pre = n * world * args.steps / (pre_ms * 1e-3)

    # e2e, this rank's own GPU: host C-ABI call, scalars from pinned host memory every step, result read back
    pinned = torch.from_numpy(np.array(ss)).pin_memory()
    pin_np = pinned.numpy()
    e2e_rank, e2e_rank_pre, e2e_cold = None, None, None
    if world == 1:
        dt, out = time_host_calls(lambda: lib.msm(hb, pin_np), args.steps)
        assert affine_of(out) == expected
        e2e_rank = n * args.steps / dt
        dt, out = time_host_calls(lambda: lib.msm(hb_pre, pin_np), args.steps)
        assert affine_of(out) == expected
        e2e_rank_pre = n * args.steps / dt
        pin_bases = torch.from_numpy(bases).pin_memory().numpy()
        dt, out = time_host_calls(lambda: lib.msm(pin_bases, pin_np), args.steps)
        assert affine_of(out) == expected
        e2e_cold = n * args.steps / dt
        del pin_bases
    e2e_ranks = None
    if world > 1:
        # per-rank host calls + NCCL all-gather + fold: the multi-process form of the same end-to-end path; the headline
        # e2e at N > 1 is the single-process call below, this one is the fallback when rank 0 cannot see every GPU
        def rank_call():
            part = torch.from_numpy(np.array(lib.msm(hb, pin_np))).to(dev)
            dist.all_gather_into_tensor(d_gather, part)
            lib.fold_g1_device(d_gather.data_ptr(), world, d_final.data_ptr(), stream.cuda_stream)
            return d_final.cpu()
        barrier()
        dt, _ = time_host_calls(rank_call, args.steps)
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ranks = n * world * args.steps / float(te.item())
    hb.free()
    hb_pre.free()
    del d_bases, d_scalars

    # e2e, N > 1: one dg_msm_g1_sharded call per step from rank 0 over the whole N * 2^logn-term input
    e2e_sp = None
    cpu_barrier()
    if single_process:
        parts = [synth_scalars(n, r) for r in range(world)]
        g_ss = np.concatenate([p[0] for p in parts])
        g_ks = np.concatenate([p[1] for p in parts])
        g_bases = gpu_bases(lib, g_ks)
        g_expected = expected_point(g_ks, g_ss)
        g_pin = torch.from_numpy(g_ss).pin_memory().numpy()
        hs = lib.ShardedBases(g_bases)
        dt, out = time_host_calls(lambda: lib.msm_sharded(hs, g_pin), args.steps)
        if affine_of(out) != g_expected:
            raise SystemExit('bench: dg_msm_g1_sharded differs from the known-dlog identity')
        e2e_sp = {'value': n * world * args.steps / dt, 'ms_per_step': 1e3 * dt / args.steps}
        hs.precompute(args.precompute_window)
        dt, out = time_host_calls(lambda: lib.msm_sharded(hs, g_pin), args.steps)
        if affine_of(out) != g_expected:
            raise SystemExit('bench: dg_msm_g1_sharded (resident table) differs from the known-dlog identity')
        e2e_sp['value_resident_table'] = n * world * args.steps / dt
        hs.free()
        del parts, g_ss, g_ks, g_bases, g_pin
    cpu_barrier()

    # ---------------------------------------------------------------- part B: sweep + one 2^24-term MSM
    sweep, strong = None, None
    big = args.strong_logn
    if big and not args.no_sweep:
        nbig = 1 << big
        g_ss, g_ks = synth_scalars(nbig, 0)              # the same global input on every rank
        ssteps = max(3, min(args.steps, 5))
        if world == 1:
            g_bases = gpu_bases(lib, g_ks)
            d_b = torch.from_numpy(g_bases).to(dev)
            d_s = torch.from_numpy(np.array(g_ss)).to(dev)
            sweep = {}
            for logn in SWEEP:
                if logn > big:
                    continue
                m = 1 << logn
                exp_m = expected_point(g_ks[:32 * m], g_ss[:32 * m])
                hp = lib.Bases(g_bases[:96 * m]).precompute(0)
                sp_, sh_ = msm_step(m, d_s, d_bases=d_b), msm_step(m, d_s, handle=hp)
                check(sp_, exp_m, 'sweep 2^%d raw bases' % logn)
                check(sh_, exp_m, 'sweep 2^%d resident table' % logn)
                ms_p = timed(sp_, ssteps, 3) / ssteps
                ms_h = timed(sh_, ssteps, 3) / ssteps
                hp.free()
                sweep[str(logn)] = {'ms': ms_p, 'muls_per_s': m / (ms_p * 1e-3), 'ms_resident_table': ms_h,
                                    'muls_per_s_resident_table': m / (ms_h * 1e-3), 'bit_exact': True}
            top = sweep[str(big)]
            strong = {'terms': nbig, 'n_gpus': 1, 'ms_per_step': top['ms'], 'value': top['muls_per_s'],
                      'ms_per_step_resident_table': top['ms_resident_table'], 'value_resident_table': top['muls_per_s_resident_table'],
                      'speedup_vs_n1': 1.0}
            pin_big = torch.from_numpy(np.array(g_ss)).pin_memory().numpy()
            hb_big = lib.Bases(g_bases)
            dt, out = time_host_calls(lambda: lib.msm(hb_big, pin_big), ssteps)
            assert affine_of(out) == expected_point(g_ks, g_ss)
            strong['e2e_ms_per_step'] = 1e3 * dt / ssteps
            strong['e2e_value'] = nbig * ssteps / dt
            hb_big.free()
            del d_b, d_s, g_bases, pin_big
        else:
            from crypto_b200 import sharding
            lo, hi = sharding.shard_range(nbig, rank, world)
            m = hi - lo
            my_ks, my_ss = g_ks[32 * lo:32 * hi], g_ss[32 * lo:32 * hi]
            my_bases = gpu_bases(lib, my_ks)
            d_b = torch.from_numpy(my_bases).to(dev)
            d_s = torch.from_numpy(np.array(my_ss)).to(dev)
            hp = lib.Bases(my_bases).precompute(0)
            exp_all = expected_point(g_ks, g_ss)
            sp_, sh_ = msm_step(m, d_s, d_bases=d_b), msm_step(m, d_s, handle=hp)
            check(sp_, exp_all, 'strong 2^%d raw bases' % big, combined=True)
            check(sh_, exp_all, 'strong 2^%d resident table' % big, combined=True)
            ms_p = timed(sp_, ssteps, 3) / ssteps
            ms_h = timed(sh_, ssteps, 3) / ssteps
            hp.free()
            del d_b, d_s
            strong = {'terms': nbig, 'n_gpus': world, 'ms_per_step': ms_p, 'value': nbig / (ms_p * 1e-3),
                      'ms_per_step_resident_table': ms_h, 'value_resident_table': nbig / (ms_h * 1e-3),
                      'result_check': 'known-dlog identity on the folded result, every rank'}
            cpu_barrier()
            if rank == 0:
                # the same MSM on this GPU alone (the other ranks idle at a host barrier): the N = 1 reference point
                g_bases = gpu_bases(lib, g_ks)
                d_b = torch.from_numpy(g_bases).to(dev)
                d_s = torch.from_numpy(np.array(g_ss)).to(dev)
                hp = lib.Bases(g_bases).precompute(0)
                s1p, s1h = msm_step(nbig, d_s, d_bases=d_b, combine=False), msm_step(nbig, d_s, handle=hp, combine=False)
                check(s1p, exp_all, 'single-GPU 2^%d raw bases' % big)
                check(s1h, exp_all, 'single-GPU 2^%d resident table' % big)
                n1_p = timed(s1p, ssteps, 3, collective=False) / ssteps
                n1_h = timed(s1h, ssteps, 3, collective=False) / ssteps
                hp.free()
                del d_b, d_s
                strong.update({'n1_ms_per_step': n1_p, 'speedup_vs_n1': n1_p / ms_p, 'n1_ms_per_step_resident_table': n1_h,
                               'speedup_vs_n1_resident_table': n1_h / ms_h,
                               'n1_note': 'the same 2^%d-term MSM on rank 0\'s GPU alone, same run, same inputs' % big})
                if single_process:
                    pin_big = torch.from_numpy(np.array(g_ss)).pin_memory().numpy()
                    hs = lib.ShardedBases(g_bases)
                    dt, out = time_host_calls(lambda: lib.msm_sharded(hs, pin_big), ssteps)
                    if affine_of(out) != exp_all:
                        raise SystemExit('bench: dg_msm_g1_sharded (2^%d) differs from the known-dlog identity' % big)
                    hb1 = lib.Bases(g_bases)
                    dt1, out1 = time_host_calls(lambda: lib.msm(hb1, pin_big), ssteps)
                    assert affine_of(out1) == exp_all
                    hb1.free()
                    hs.free()
                    strong.update({'e2e_ms_per_step': 1e3 * dt / ssteps, 'e2e_value': nbig * ssteps / dt,
                                   'e2e_n1_ms_per_step': 1e3 * dt1 / ssteps, 'e2e_speedup_vs_n1': dt1 / dt,
                                   'e2e_note': 'dg_msm_g1_sharded: one host call from one thread, scalars scattered from pinned host '
                                               'memory over %d PCIe links, partials folded out of peer memory; n1 = dg_msm_g1' % world})
                    del pin_big
                del g_bases
            cpu_barrier()
        del g_ss, g_ks
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return

    # ---------------------------------------------------------------- the line
    peaks = load_json(os.path.join(ROOT, 'MEASURED_PEAKS.json')) or {}
    hbm_peak = peaks.get('hbm_gbs')
    peak_src = 'measured (MEASURED_PEAKS.json)'
    if not hbm_peak:
        hbm_peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    alg_bytes = 128.0 * n                                   # 32 B scalar + 96 B affine base per term
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms else None
    ip = load_json(os.path.join(ROOT, 'profiles', 'int_peak_r01.json')) or {}
    imad_peak = (ip.get('imad_lo') or {}).get('ops_per_s')
    from crypto_b200 import msm as msm_mirror
    c_plain, rounds = lib.msm_plan(n)
    ndig = msm_mirror.msm_digits_per_scalar(c_plain, plain_bases=True)              # GLV: 2 x ceil(128 / c) digits per scalar
    # The dominant kernel of the `value` path: with batch-affine rounds it is round 0 (k_affine_round<Fp, gather>), which
    # visits every (scalar digit, base) entry once and performs half of them as affine additions; without rounds k_accumulate.
    if rounds:
        dom_kernel = 'k_affine_round<Fp, gather> (round 0 of %d, raw bases + GLV, c = %d, %d digits per scalar)' % (rounds, c_plain, ndig)
        ncu = load_json(os.path.join(ROOT, 'profiles', 'ncu_affine_round_plain_r02.json')) or \
            load_json(os.path.join(ROOT, 'profiles', 'ncu_affine_round.json')) or {}
        adds, mults_per_add = n * ndig / 2.0, 6.0          # 5M + 1S per affine addition incl. the shared inversion
    else:
        dom_kernel = 'k_accumulate<Fp>'
        ncu = load_json(os.path.join(ROOT, 'profiles', 'ncu_accumulate.json')) or {}
        adds, mults_per_add = float(n * ndig), 10.0        # 8M + 2S per XYZZ mixed addition
    cfg = bench_config(args.logn, world)
    if world == 1:
        e2e = {'value': e2e_rank, 'unit': UNIT, 'h2d_bytes_per_step': 32 * n, 'd2h_bytes_per_step': 144,
               'note': 'dg_msm_g1 host call; scalars from pinned host memory every step; raw bases resident behind a plain handle',
               'value_resident_table': e2e_rank_pre, 'value_cold_bases': e2e_cold,
               'cold_bases_note': 'bases (96 B/term) also copied from pinned host memory every step'}
    elif e2e_sp:
        e2e = {'value': e2e_sp['value'], 'unit': UNIT, 'h2d_bytes_per_step': 32 * n * world, 'd2h_bytes_per_step': 144,
               'ms_per_step': e2e_sp['ms_per_step'], 'value_resident_table': e2e_sp['value_resident_table'],
               'note': 'ONE dg_msm_g1_sharded call per step from one host thread of rank 0 driving all %d GPUs: scalars scattered from '
                       'pinned host memory by one worker thread per device, raw bases resident per device, partial results folded on '
                       'device 0 out of peer memory (NVLink), 144-byte result read back' % world}
    else:
        e2e = {'value': e2e_ranks, 'unit': UNIT, 'h2d_bytes_per_step': 32 * n * world, 'd2h_bytes_per_step': 144 * world,
               'note': 'rank 0 does not see every GPU of the job, so the single-process sharded call was not measured: every rank '
                       'calls dg_msm_g1 on its shard (scalars from pinned host memory), NCCL all-gather of the partial results, fold, read back'}
    if world > 1 and e2e_ranks:
        e2e['value_one_process_per_gpu'] = e2e_ranks
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': plain_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u32', 'data': 'synthetic',
        'value_resident_table': value_pre, 'ms_per_step_resident_table': pre_ms / args.steps,
        'config': cfg,
        'notes': {'value': 'raw 96-byte bases and scalars resident in HBM, nothing precomputed (like for like with msm_bigint)',
                  'value_resident_table': 'bases behind a handle carrying the 2^(ck)-multiples table built once at upload '
                                          '(dg_bases_precompute: proving keys / signature parameters are fixed across calls)',
                  'parallelism': 'base-range shards x%d + all-gather/fold' % world,
                  'l2': 'flushed between timed iterations (256 MiB fill)', 'result_check': 'known-dlog identity, bit-exact, before timing'},
        'clocks': clocks,
        'e2e': e2e,
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': dom_kernel, 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': (achieved / hbm_peak) if achieved else None, 'traffic': ncu.get('dram_bytes_per_launch'),
                     'peak_source': peak_src, 'kernel_ms': dom_ms, 'launches_timed': dom_cnt,
                     'algorithmic_bytes_per_launch': alg_bytes,
                     'note': 'integer-issue bound, not HBM bound (SURVEY 8d): see int_roofline'},
    }
    if sweep:
        line['sweep'] = sweep
    if strong:
        line['strong_2p%d' % big] = strong
    if imad_peak and dom_ms:
        # 32x32 multiply-accumulates the dominant launch algorithmically needs: its additions x Fp mults per
        # addition x 2*12^2 MACs per Montgomery mult
        macs = adds * mults_per_add * 288.0
        line['int_roofline'] = {'bound': 'imad', 'achieved': macs / (dom_ms * 1e-3), 'peak': imad_peak, 'unit': 'MAC/s',
                                'frac': macs / (dom_ms * 1e-3) / imad_peak,
                                'peak_source': 'profiles/int_peak_r01.json imad_lo (measured on this pool)'}
        # The multiplier needs 64-bit products: IMAD.WIDE / IMAD.HI issue at half the IMAD rate (profiles/int_peak_r02.json,
        # operands that ptxas cannot hoist), so the ceiling for 32x32->64 multiply-accumulates is the product-pair figure.
        ip2 = load_json(os.path.join(ROOT, 'profiles', 'int_peak_r02.json')) or {}
        wide_peak = (ip2.get('split_mul_lo_hi_addc_products') or {}).get('ops_per_s')
        if wide_peak:
            line['int_roofline']['peak_wide_mac'] = wide_peak
            line['int_roofline']['frac_wide_mac'] = macs / (dom_ms * 1e-3) / wide_peak
            line['int_roofline']['peak_wide_mac_source'] = ('profiles/int_peak_r02.json split_mul_lo_hi_addc_products: 32x32->64 '
                                                            'products per second (IMAD + IMAD.HI + carry adds), measured on this pool')
    fp = load_json(os.path.join(ROOT, 'profiles', 'fpmul_peak_r01.json')) or {}
    mult_peak = fp.get('fp_mul_12x32_carry_chain_mults_per_s')
    if mult_peak and dom_ms:
        mults = adds * mults_per_add
        line['mult_roofline'] = {'bound': 'fp-multiplier issue', 'achieved': mults / (dom_ms * 1e-3), 'peak': mult_peak,
                                 'unit': 'Fp mult/s', 'frac': mults / (dom_ms * 1e-3) / mult_peak,
                                 'peak_source': 'profiles/fpmul_peak_r01.json (tools/fpmul_bench.cu, measured on this pool)'}
    if world == 1 and not args.no_cpu:
        from oracle import cref
        cores = host_cores()
        cref.set_threads(cores)
        dt, out = time_cpu_msm(bases, ss, n, 2)
        assert affine_of(out) == expected
        line['cpu_baseline'] = {'value': n / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': 'best of 2 full 2^%d-term MSMs, oracle C restatement of ark-ec 0.4 '
                                          'msm_bigint_wnaf (OpenMP over windows)' % args.logn}
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


_REAL_STDOUT = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--logn', type=int, default=20, help='log2 of the terms per GPU')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-sweep', action='store_true', help='skip the 2^16..2^24 sweep and the 2^24-term strong-scaling MSM')
    ap.add_argument('--no-single-process', action='store_true', help='N > 1: skip the dg_msm_g1_sharded legs on rank 0')
    ap.add_argument('--strong-logn', type=int, default=24, help='log2 of the GLOBAL term count of the strong-scaling MSM')
    ap.add_argument('--clock-sample-ms', type=int, default=200, help='nvidia-smi polling period during the timed regions')
    ap.add_argument('--precompute-window', type=int, default=0, help='window bits of the resident table (0 = default)')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world == 1:
        use_all_host_cores()                 # the cpu_baseline leg; before anything loads libgomp
    if world > 1:
        # NCCL prints its version banner on stdout; keep stdout for the one JSON line (everything else -> stderr)
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)
        import torch
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('GLOO_SOCKET_IFNAME', 'lo')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
