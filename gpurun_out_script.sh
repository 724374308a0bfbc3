mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_neighbours.py -m gpu -x -q -k hilbert ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --config h1 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_h1.json 2> gpurun_out/bench_h1.err; tail -2 gpurun_out/bench_h1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_h1.json').read().strip().splitlines()[-1]); print('h1', round(d['value']), d['ms_per_step'], d['roofline']['frac'])
PY
