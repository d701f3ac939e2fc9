#!/usr/bin/env python3
"""bench.py -- BLS12-381 G1 MSM throughput (BASELINE.json metric) on N x B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--logn 20] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...          (one rank per GPU, NCCL)

A "step" is one G1 MSM over synthetic seeded scalars/bases (bases k_i*G with known k_i, scalars
uniform in [0, r)).  N = 1: 2^20 terms (the size the metric is quoted on).  N > 1: weak scaling,
every rank owns a contiguous 2^20-term base range of one N*2^20-term MSM; partial results are
all-gathered (NCCL, 144 B per rank) and folded on the GPU under the group law.
`value`   : terms/s with scalars and bases already resident in HBM (dg_msm_g1_handle_device).
            Bases live behind a handle as in the reference's workloads (proving keys / signature
            parameters are fixed across calls) with the 2^(20k)-multiples table built
            once at upload (dg_bases_precompute, 15 x the base memory at 2^20).  `value_plain_bases` is the
            same MSM through dg_msm_g1_device on the raw 96-byte bases with nothing precomputed.
`e2e`     : terms/s through the host C-ABI call dg_msm_g1 with the scalars in pinned host memory
            copied every step and the 144-byte result read back every step (same handle);
            `e2e_plain_bases` likewise without the precomputed table.
--total-logn T switches to strong scaling: one 2^T-term MSM split over the ranks.
`roofline`: the dominant kernel (round 0 of the batch-affine stage, else k_accumulate) against the measured HBM peak, algorithmic bytes
            128 B/term; `int_roofline` is the integer-pipe reading of the same launch.
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


def synth_inputs(n, rank=0):
    """Seeded inputs: scalars uniform in [0, r); bases k_i * G with known k_i.
    Generated with the oracle's fixed-base helper (input synthesis, not the measured path)."""
    from oracle import cref
    seed = (0xD0C4C0DE ^ n) + 7919 * rank
    scalars = cref.random_scalars(n, seed)
    ks = cref.random_scalars(n, seed + 1)
    bases = cref.g1_generator_muls(ks)
    return bases, scalars, ks


def known_dlog_expected(ks, scalars):
    from oracle import cref
    r = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    k = np.asarray(ks, dtype=np.uint8).reshape(-1, 32)
    s = np.asarray(scalars, dtype=np.uint8).reshape(-1, 32)
    tot = 0
    for a, b in zip(k, s):
        tot += int.from_bytes(bytes(a), 'little') * int.from_bytes(bytes(b), 'little')
    return bytes(cref.g1_generator_muls(np.frombuffer((tot % r).to_bytes(32, 'little'), dtype=np.uint8)))


class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
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
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower() == 'active':
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
    msm_bigint_wnaf, OpenMP over windows like rayon) on the host cores, same config/metric."""
    if rank != 0:
        return
    n = 1 << (args.total_logn if args.total_logn else args.logn)
    bases, scalars, ks = synth_inputs(n)
    cores = host_cores()
    os.environ.setdefault('OMP_NUM_THREADS', str(cores))
    for _ in range(min(args.warmup, 1)):
        time_cpu_msm(bases, scalars, n, 1)
    times = []
    for _ in range(args.steps):
        dt, _ = time_cpu_msm(bases, scalars, n, 1)
        times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': {'workload': 'bls12-381 g1 msm, 2^%d random scalars/bases' % args.logn, 'terms': n},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d x full 2^%d-term MSM (C restatement of ark-ec 0.4 msm_bigint_wnaf, OpenMP over '
                                   'windows; the Rust reference cannot be built in this image)' % (len(times), args.logn)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def load_profile_json(name):
    path = os.path.join(ROOT, 'profiles', name)
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
    lib.init(local_rank)
    strong = args.total_logn > 0
    if strong:
        n = (1 << args.total_logn) // world  # strong scaling: fixed global MSM split by base range
    else:
        n = 1 << args.logn                   # weak scaling: terms per rank
    bases, scalars, ks = synth_inputs(n, rank)
    d_bases = torch.from_numpy(bases).to(dev)
    d_scalars = torch.from_numpy(scalars).to(dev)
    d_out = torch.zeros(144, dtype=torch.uint8, device=dev)
    d_gather = torch.zeros(144 * world, dtype=torch.uint8, device=dev)
    d_final = torch.zeros(144, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    hb = lib.Bases(bases)
    hb_pre = lib.Bases(bases).precompute(args.precompute_window)
    mode = {'pre': True}

    def step():
        if mode['pre']:
            lib.msm_handle_device(hb_pre, d_scalars.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)
        else:
            lib.msm_device(d_bases.data_ptr(), d_scalars.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(d_gather, d_out)
            lib.fold_g1_device(d_gather.data_ptr(), world, d_final.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness gate before timing: known-discrete-log identity on this rank's shard, both paths
    from oracle import cref
    expected = known_dlog_expected(ks, scalars)
    for pre in (False, True):
        mode['pre'] = pre
        for _ in range(args.warmup):
            flush.fill_(1)
            step()
        barrier()
        got = bytes(cref.normalize_batch_g1(d_out.cpu().numpy()))
        if got != expected:
            raise SystemExit('bench: GPU MSM result differs from the known-dlog identity (precomputed=%s)' % pre)

    def timed(pre):
        mode['pre'] = pre
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for e0, e1 in evs:
            flush.fill_(1)                   # L2 flush between timed iterations (outside the event pair)
            e0.record(stream)
            step()
            e1.record(stream)
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clocks are sampled from here to the end of the e2e loops (every timed region; ~100 ms steps of nvidia-smi)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)                      # let nvidia-smi come up before the first timed region
    plain_ms = timed(False)

    lib.prof_enable(True)
    lib.prof_read_accumulate()
    launches0 = lib.launch_count()
    total_ms = timed(True)
    launches = lib.launch_count() - launches0
    acc_ms, acc_cnt = lib.prof_read_accumulate()
    lib.prof_enable(False)
    value = n * world * args.steps / (total_ms * 1e-3)
    value_plain = n * world * args.steps / (plain_ms * 1e-3)

    # ---- e2e: host C-ABI call, scalars from pinned host memory every step, result read back ----
    pinned = torch.from_numpy(scalars.copy()).pin_memory()
    pin_np = pinned.numpy()

    def e2e(handle):
        for _ in range(2):
            lib.msm(handle, pin_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_host = lib.msm(handle, pin_np)
            if world > 1:
                part = torch.from_numpy(np.array(out_host)).to(dev)
                dist.all_gather_into_tensor(d_gather, part)
                lib.fold_g1_device(d_gather.data_ptr(), world, d_final.data_ptr(), stream.cuda_stream)
                d_final.cpu()
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return n * world * args.steps / float(te.item())

    e2e_plain = e2e(hb)
    e2e_value = e2e(hb_pre)
    clocks = sampler.stop() if rank == 0 else None
    hb.free()
    hb_pre.free()

    if rank != 0:
        return
    peaks = load_profile_json('../MEASURED_PEAKS.json') or {}
    hbm_peak = peaks.get('hbm_gbs')
    peak_src = 'measured (MEASURED_PEAKS.json)'
    if not hbm_peak:
        hbm_peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    alg_bytes = 128.0 * n                                   # 32 B scalar + 96 B affine base per term
    achieved = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms else None
    ip = load_profile_json('int_peak_r01.json') or {}
    imad_peak = (ip.get('imad_lo') or {}).get('ops_per_s')
    pre_c = args.precompute_window if args.precompute_window else (20 if n >= (1 << 22) else 17)   # dg_bases_precompute default
    nwin = (254 + pre_c - 1) // pre_c + (1 if 254 % pre_c == 0 else 0)        # msm_ndigits
    _, rounds = lib.msm_plan(n, precomputed_c=pre_c)
    # The dominant kernel: with batch-affine rounds it is round 0 (k_affine_round<Fp, gather>), which visits every
    # (scalar digit, base) entry once and performs half of them as affine additions; without rounds k_accumulate.
    if rounds:
        dom_kernel = 'k_affine_round<Fp, gather> (round 0 of %d)' % rounds
        ncu = load_profile_json('ncu_affine_round.json') or {}
        adds, mults_per_add = n * nwin / 2.0, 6.0          # 5M + 1S per affine addition incl. the shared inversion
    else:
        dom_kernel = 'k_accumulate<Fp>'
        ncu = load_profile_json('ncu_accumulate.json') or {}
        adds, mults_per_add = float(n * nwin), 10.0        # 8M + 2S per XYZZ mixed addition
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak',
        'vs_baseline': None, 'dtype': 'u32', 'data': 'synthetic',
        'value_plain_bases': value_plain, 'ms_per_step_plain_bases': plain_ms / args.steps,
        'config': {'workload': 'bls12-381 g1 msm, 2^%.3g random scalars/bases per GPU' % np.log2(n), 'terms_per_gpu': n,
                   'global_terms': n * world, 'parallelism': 'base-range shards x%d + all-gather/fold' % world,
                   'bases': 'resident behind a handle, 2^(%d k)-multiples table (%d rows) built once at upload; '
                            'value_plain_bases = raw bases, nothing precomputed' % (pre_c, nwin),
                   'l2': 'flushed between timed iterations (256 MiB fill)', 'result_check': 'known-dlog identity, bit-exact'},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 32 * n, 'd2h_bytes_per_step': 144,
                'note': 'dg_msm_g1 host call; scalars from pinned host memory every step; bases resident (handle)'},
        'e2e_plain_bases': e2e_plain,
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': dom_kernel, 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': (achieved / hbm_peak) if achieved else None, 'traffic': ncu.get('dram_bytes_per_launch'),
                     'peak_source': peak_src, 'kernel_ms': acc_ms, 'launches_timed': acc_cnt,
                     'algorithmic_bytes_per_launch': alg_bytes,
                     'note': 'integer-issue bound, not HBM bound (SURVEY 8d): see int_roofline'},
    }
    if imad_peak and acc_ms and nwin:
        # 32x32 multiply-accumulates the dominant launch algorithmically needs: its additions x Fp mults per
        # addition x 2*12^2 MACs per Montgomery mult
        macs = adds * mults_per_add * 288.0
        line['int_roofline'] = {'bound': 'imad', 'achieved': macs / (acc_ms * 1e-3), 'peak': imad_peak, 'unit': 'MAC/s',
                                'frac': macs / (acc_ms * 1e-3) / imad_peak,
                                'peak_source': 'profiles/int_peak_r01.json imad_lo (measured on this pool)'}
    fp = load_profile_json('fpmul_peak_r01.json') or {}
    mult_peak = fp.get('fp_mul_12x32_carry_chain_mults_per_s')
    if mult_peak and acc_ms:
        mults = adds * mults_per_add
        line['mult_roofline'] = {'bound': 'fp-multiplier issue', 'achieved': mults / (acc_ms * 1e-3), 'peak': mult_peak,
                                 'unit': 'Fp mult/s', 'frac': mults / (acc_ms * 1e-3) / mult_peak,
                                 'peak_source': 'profiles/fpmul_peak_r01.json (tools/fpmul_bench.cu, measured on this pool)'}
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        os.environ.setdefault('OMP_NUM_THREADS', str(cores))
        dt, _ = time_cpu_msm(bases, scalars, n, 2)
        line['cpu_baseline'] = {'value': n / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': 'best of 2 full 2^%d-term MSMs, oracle C restatement of ark-ec 0.4 '
                                          'msm_bigint_wnaf (OpenMP over windows)' % args.logn}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--logn', type=int, default=20, help='log2 of the terms per GPU')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--total-logn', type=int, default=0, help='strong scaling: log2 of the GLOBAL term count')
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
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
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
