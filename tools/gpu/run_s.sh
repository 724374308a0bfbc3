#!/bin/bash
# GPU call S: polyphase kernel experiments on config 5 (tune bits: 1 evict-first gather, 2 no prefetch, 4 no resident spectrum rows; groups cap)
mkdir -p gpurun_out
rm -f gpurun_out/s_c5_variants.txt gpurun_out/s_err.txt
for v in "4 0 0" "4 4 0" "4 16 0" "4 20 0" "2 4 0" "2 20 0" "2 16 0"; do
  set -- $v
  echo "== C=$1 tune=$2 groups=$3" | tee -a gpurun_out/s_c5_variants.txt
  RRC_FFTFILT_POLY_C=$1 RRC_FFTFILT_POLY_TUNE=$2 RRC_FFTFILT_POLY_GROUPS=$3 timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>>gpurun_out/s_err.txt | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('roofline', {}).get('frac'))
" | tee -a gpurun_out/s_c5_variants.txt
done
tail -3 gpurun_out/s_err.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "polyphase" 2>&1 | tail -3 | tee gpurun_out/s_pytest.txt
RRC_FFTFILT_TRACE=1 RRC_FFTFILT_POLY_C=4 timeout 300 python bench.py --config c5 --steps 2 --warmup 1 --headline-only --no-e2e --no-cpu --sustain 0 2>&1 >/dev/null | grep -A15 "iter [456]:" > gpurun_out/s_trace.txt
