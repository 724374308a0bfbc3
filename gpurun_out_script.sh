mkdir -p gpurun_out
fails=0
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/flakeall_$i.log 2>&1 || { fails=$((fails+1)); grep -n "FAILED" gpurun_out/flakeall_$i.log | head -2; }; tail -1 gpurun_out/flakeall_$i.log; done
echo "full suite x2: $fails failures"
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_default_final.json').read().strip().splitlines()[-1]); print('default', d['config']['name'], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['e2e']['value']), d['gpu_launches'])"
