mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_all.log 2>&1; tail -2 gpurun_out/pytest_gpu_all.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_default_final.json').read().strip().splitlines()[-1]); print('default', d['config']['name'], round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), round(d['e2e']['value']), d['gpu_launches'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
for cfg in c1 c3; do timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_x.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1]); print('$cfg', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), d['roofline']['kernel'][:30])"; done
