#!/bin/bash
# GPU call F: packed kernel with FP-turn ping-pong (variant 38) vs 37 vs 36: parity, timing, per-phase trace.
mkdir -p gpurun_out
RRC_FFTFILT_VARIANT=39 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fftfilt or config2 or halo" > gpurun_out/f_parity38.log 2>&1; echo "parity38 rc=$?"; tail -2 gpurun_out/f_parity38.log
for v in 36 37 39; do
  RRC_FFTFILT_VARIANT=$v timeout 300 python bench.py --config c2 --steps 30 --warmup 5 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/f_c2_v$v.json 2>gpurun_out/f_c2_v$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/f_c2_v$v.json').read().strip().splitlines()[-1]); print('c2 variant $v', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
done
for v in 39; do
  RRC_FFTFILT_TRACE=1 RRC_FFTFILT_VARIANT=$v timeout 300 python bench.py --config c2 --n 67108864 --steps 1 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > /dev/null 2> gpurun_out/f_trace_v$v.txt
done
