#!/bin/bash
# GPU call W: final tree — full GPU suite, smoke, config-5 line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/w_pytest_full.txt 2>&1; grep -E "passed|failed" gpurun_out/w_pytest_full.txt | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/w_smoke.txt
timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --headline-only --no-cpu --sustain 0 2>gpurun_out/w_err.txt | tail -1 > gpurun_out/w_bench_c5.json; python -c "
import json; d=json.loads(open('gpurun_out/w_bench_c5.json').read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"
