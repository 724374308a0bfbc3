#!/bin/bash
# GPU call D: packed FftFilter kernel (variant 37) — parity + timing against variant 36.
mkdir -p gpurun_out
RRC_FFTFILT_VARIANT=37 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_blocks_gpu.py tests/test_baseline_size.py -m gpu -x -q -k "fftfilt or fft_filter or FftFilter or fftfilter or config2 or config5 or halo" > gpurun_out/d_parity37.log 2>&1; echo "parity37 rc=$?"; tail -3 gpurun_out/d_parity37.log
timeout 900 python -m pytest tests/test_sources.py -m gpu -x -q > gpurun_out/d_sources.log 2>&1; echo "sources rc=$?"; tail -2 gpurun_out/d_sources.log
for v in 36 37; do for t in 1 0; do
  RRC_FFTFILT_VARIANT=$v RRC_FFTFILT_TUNE=$t timeout 300 python bench.py --config c2 --steps 30 --warmup 5 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/d_c2_v$v_t$t.json 2>gpurun_out/d_c2_v${v}_t$t.err
  python -c "
import json; d=json.loads(open('gpurun_out/d_c2_v$v_t$t.json').read().strip().splitlines()[-1]); print('c2 variant $v tune $t', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
done; done
