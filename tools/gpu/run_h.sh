#!/bin/bash
mkdir -p gpurun_out
RRC_FFTFILT_TRACE=1 timeout 300 python bench.py --config c5 --n 268435456 --steps 1 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/h_c5.json 2> gpurun_out/h_trace_c5.txt
grep -A11 "block iter 4" gpurun_out/h_trace_c5.txt
for c in c2 c5; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"; done
