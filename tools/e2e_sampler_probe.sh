# e2e per call (ms) of bench.py for different step counts / sampler periods: bash tools/e2e_sampler_probe.sh
for st in 3 5 5 10 20; do
timeout 200 python bench.py --no-sweep --no-cpu --steps $st 2>/dev/null > /tmp/b.json
python - $st <<'PY'
import sys, json
d = json.load(open('/tmp/b.json'))
print('steps', sys.argv[1], 'device ms', round(d["ms_per_step"], 3), 'e2e ms', round(2**20 / d["e2e"]["value"] * 1e3, 3))
PY
done
