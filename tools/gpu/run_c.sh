#!/bin/bash
# GPU call C: full GPU suite + timing of a config-2 shard-sized run (2^25) on one GPU.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/c_suite.log
for n in 33554432 268435456; do
  timeout 300 python bench.py --config c2 --n $n --steps 50 --warmup 5 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/c_c2_n$n.json 2>gpurun_out/c_c2_n$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/c_c2_n$n.json').read().strip().splitlines()[-1]); print('c2 n=$n', round(d['value']), round(d['ms_per_step'],4), d['gpu_launches'])"
done
