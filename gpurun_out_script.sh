mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_neighbours.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --config e4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_e4.json 2> gpurun_out/bench_e4.err; tail -2 gpurun_out/bench_e4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_e4.json').read().strip().splitlines()[-1]); print('e4', round(d['value']), d['ms_per_step'], d['roofline']['frac'])
PY
