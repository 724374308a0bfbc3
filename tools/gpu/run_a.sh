#!/bin/bash
# GPU call A: new BASELINE-size tests, the whole GPU suite, smoke, default bench (all configs).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
nproc; cat /sys/devices/system/node/online 2>/dev/null
timeout 1500 python -m pytest tests/test_baseline_size.py -m gpu -x -q -s > gpurun_out/a_baseline_size.log 2>&1; echo "baseline-size rc=$?"; tail -5 gpurun_out/a_baseline_size.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_baseline_size.py > gpurun_out/a_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/a_suite.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/a_smoke.log 2>&1; tail -1 gpurun_out/a_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -4 gpurun_out/a_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/a_bench.json').read().strip().splitlines()[-1])
    print('c2', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), 'e2e', d['e2e'] and round(d['e2e']['value']), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],4))
    for k,v in d.get('configs',{}).items():
        print(k, v.get('error') or (round(v['value']), round(v['ms_per_step'],4), round(v['roofline']['frac'],3), v.get('e2e') and round(v['e2e']['value'])))
except Exception as e:
    print('bench parse failed', e)
PY
