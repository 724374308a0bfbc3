#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_baseline_size.py tests/test_ingest.py -m gpu -x -q -k "decim or config5 or halo or fold or u8" > gpurun_out/k_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/k_parity.log
for c in c5 c5u8 c2; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"; done
RRC_FFTFILT_TRACE=1 timeout 300 python bench.py --config c5 --n 268435456 --steps 1 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > /dev/null 2> gpurun_out/k_trace_c5.txt
grep -A11 "block iter 4" gpurun_out/k_trace_c5.txt
