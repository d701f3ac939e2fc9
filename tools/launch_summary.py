"""Per-kernel device time of the LAST step in an ncu launch list (gpu__time_duration.sum CSV):
   python tools/launch_summary.py gpurun_out/launches.csv 'command' > profiles/launches_xxx_summary.json
A step is delimited by k_digits<0> launches; times are cold-cache and serialised: compare shares."""
import csv, json, sys, collections

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
data = [(r[ki], float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)) for r in rows[1:]]
# a step starts at a k_digits<0> launch that does not directly follow another one (the host path counts in four chunks)
starts = [i for i, (k, _) in enumerate(data) if 'k_digits<0>' in k and (i == 0 or 'k_digits<0>' not in data[i - 1][0])]
# raw bases: the GLV expansion of the bases is the first kernel of the step
starts = [i - 1 if i > 0 and 'k_glv_expand' in data[i - 1][0] else i for i in starts]
steps = []
for a, b in zip(starts, starts[1:] + [len(data)]):
    steps.append(data[a:b])


def short(k):
    k = k.replace('void ', '').replace('dg::', '')
    return k.split('(')[0]


def summarise(step):
    agg = collections.OrderedDict()
    for k, us in step:
        if 'k_digits<0>' not in k and not any(s in k for s in ('k_', )):
            continue
        agg[short(k)] = agg.get(short(k), 0.0) + us
    tot = sum(agg.values())
    return {'total_us': round(tot, 1), 'kernels': [{'kernel': k, 'us_per_step': round(v, 1), 'share': round(v / tot, 4)} for k, v in agg.items()]}


# group the steps by their kernel signature (the bench runs a plain-bases and a precomputed variant)
out = {'command': sys.argv[2] if len(sys.argv) > 2 else '', 'note': 'cold-cache serialised per-launch times: compare shares, not absolutes',
       'steps_seen': len(steps)}
seen = {}
for st in steps:
    # drop trailing foreign kernels (torch fills etc.) that follow the window combine
    idx = [i for i, (k, _) in enumerate(st) if 'k_window_combine' in k or 'k_set_jac_inf' in k]
    if idx:
        st = st[:idx[-1] + 1]
    sig = tuple(sorted(set(short(k) for k, _ in st)))
    seen[sig] = st
for n, (sig, st) in enumerate(seen.items()):
    out['variant_%d' % n] = summarise(st)
print(json.dumps(out, indent=1))
