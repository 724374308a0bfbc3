#!/bin/bash
# GPU call G: staggered packed kernel (variant 39): step sweep, parity, trace.
mkdir -p gpurun_out
RRC_FFTFILT_VARIANT=39 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fftfilt or config2 or halo" > gpurun_out/g_parity39.log 2>&1; echo "parity39 rc=$?"; tail -2 gpurun_out/g_parity39.log
for step in 32 64 96 128 192 256; do
  RRC_FFTFILT_TUNE=$((step*65536+1)) RRC_FFTFILT_VARIANT=39 timeout 300 python bench.py --config c2 --steps 30 --warmup 5 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/g_c2_s$step.json 2>gpurun_out/g_c2_s$step.err
  python -c "
import json; d=json.loads(open('gpurun_out/g_c2_s$step.json').read().strip().splitlines()[-1]); print('c2 variant 39 step $step', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
done
RRC_FFTFILT_TRACE=1 RRC_FFTFILT_VARIANT=39 timeout 300 python bench.py --config c2 --n 67108864 --steps 1 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > /dev/null 2> gpurun_out/g_trace_v39.txt
