mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for c in a12; do timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -2 gpurun_out/bench_$c.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$c.json')); print('$c', round(d['ms_per_step'],4), 'ms', round(d['value']), 'Msps frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'] and round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1))
except Exception as e: print('$c failed', e)
PY
done
