#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_bench_default_n1_v3.json 2> gpurun_out/o_bench.err; tail -4 gpurun_out/o_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_default_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/o_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_n1_v3.json 2>> gpurun_out/o_bench.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default_n1_v3.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['gpu_launches'], d['clocks'])
for k,v in d.get('configs',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('roofline',{}).get('frac'), (v.get('roofline',{}).get('kernel') or '')[:40], v.get('e2e',{}).get('value'))
PY
