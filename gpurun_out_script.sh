mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ingest.py -m gpu -x -q 2>&1 | tail -3
for c in c3u8 c3; do timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -2 gpurun_out/bench_$c.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$c.json')); print('$c', round(d['ms_per_step'],4), 'ms', round(d['value']), 'Msps frac', round(d['roofline']['frac'],3))
except Exception as e: print('$c failed', e)
PY
done
