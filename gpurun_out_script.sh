mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_neighbours.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for c in h1 e1 e2 e3 e4; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --cpu-budget 5 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -2 gpurun_out/bench_$c.err; done
python - <<PY
import json
for c in ('h1','e1','e2','e3','e4'):
    try:
        d=json.loads(open(f'gpurun_out/bench_{c}.json').read().strip().splitlines()[-1]); print(c, round(d['value']), d['ms_per_step'], d['roofline']['frac'], (d.get('e2e') or {}).get('value'), d['cpu_baseline']['value'], d['clocks'])
    except Exception as e: print(c, 'failed', e)
PY
