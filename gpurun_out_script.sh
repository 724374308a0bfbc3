mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_all.log 2>&1; tail -2 gpurun_out/pytest_gpu_all.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --config c1 --steps 20 --warmup 3 > gpurun_out/bench_c1_v12.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_c1_v12.json').read().strip().splitlines()[-1]); print('c1', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['e2e']['value']))"
