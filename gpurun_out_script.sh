mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_v11.json 2> gpurun_out/bench_c2_v11.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2_v11.json').read().strip().splitlines()[-1]); print('c2', round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 300 python bench.py --config c1 --steps 20 --warmup 3 > gpurun_out/bench_c1_v11.json 2> gpurun_out/bench_c1_v11.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c1_v11.json').read().strip().splitlines()[-1]); print('c1', round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), d['e2e']['value'])"
timeout 300 python bench.py --config c3 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c3_v11.json 2> gpurun_out/bench_c3_v11.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_v11.json').read().strip().splitlines()[-1]); print('c3', round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), d['roofline']['kernel'][:40])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_tc1_kernel -s 3 -c 1 -f -o gpurun_out/c1_tc_v11 python bench.py --config c1 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_c1_tc.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_c1_v11.csv python bench.py --config c1 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
