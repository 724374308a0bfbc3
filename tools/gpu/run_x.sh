#!/bin/bash
# GPU call X: FftFilter variant 41 (spectrum AND phase-A twiddle powers in tensor memory) against variant 40 on config 2
mkdir -p gpurun_out
rm -f gpurun_out/x_variants.txt gpurun_out/x_err.txt
RRC_FFTFILT_VARIANT=42 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fftfilt and not polyphase and not fold" 2>&1 | tail -3 | tee gpurun_out/x_pytest.txt
for v in 41 42 41 42; do
  echo "== variant $v" | tee -a gpurun_out/x_variants.txt
  RRC_FFTFILT_VARIANT=$v timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>>gpurun_out/x_err.txt | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('roofline', {}).get('frac'))
" | tee -a gpurun_out/x_variants.txt
done
tail -3 gpurun_out/x_err.txt
