"""Summarise one kernel of an ncu report (read here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep 'command line that was profiled' 'workload text' > profiles/ncu_xxx.json
Takes the FIRST profiled launch in the report."""
import csv, io, json, subprocess, sys

rep, cmd, workload = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
u = dict(zip(hdr, units))


def f(name, scale=1.0):
    v = m.get(name)
    return None if v in (None, '') else float(v.replace(',', '')) * scale


def to_bytes(name):
    v, unit = f(name), u.get(name, '')
    if v is None:
        return None
    return int(v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1))


def to_ms(name):
    v, unit = f(name), u.get(name, '')
    return None if v is None else v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1)


stalls = {}
for k, v in m.items():
    if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
        stalls[k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = round(float(v), 3)
stalls = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
out = {
    'kernel': m.get('Kernel Name'), 'command': cmd, 'workload': workload,
    'gpu__time_duration_ms': to_ms('gpu__time_duration.sum'),
    'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_launch': (rd or 0) + (wr or 0),
    'gpu__dram_throughput_pct': f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    'launch__registers_per_thread': f('launch__registers_per_thread'),
    'launch__grid_size': f('launch__grid_size'), 'launch__block_size': f('launch__block_size'),
    'occupancy_limit_registers_blocks_per_sm': f('launch__occupancy_limit_registers'),
    'sm__warps_active_pct': f('sm__warps_active.avg.pct_of_peak_sustained_active'),
    'smsp__issue_active_pct': f('smsp__issue_active.avg.pct_of_peak_sustained_active'),
    'smsp__inst_executed_sum': f('smsp__inst_executed.sum'),
    'l2_sector_hit_rate_pct': f('lts__t_sector_hit_rate.pct'),
    'stall_breakdown_warps_per_issue': stalls,
}
print(json.dumps(out, indent=1))
