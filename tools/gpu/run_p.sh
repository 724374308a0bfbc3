#!/bin/bash
# GPU call P: polyphase decimating FftFilter kernel (fftfilt_poly.cu): parity tests, then config 5 timed per variant
mkdir -p gpurun_out
rm -f gpurun_out/p_c5_variants.txt gpurun_out/p_trace.txt gpurun_out/p_err.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "polyphase or decim or fold" 2>&1 | tail -15 | tee gpurun_out/p_pytest.txt
for v in "4 0" "4 2" "4 1" "2 0" "1 0"; do
  set -- $v
  echo "== C=$1 tune=$2" | tee -a gpurun_out/p_c5_variants.txt
  RRC_FFTFILT_POLY_C=$1 RRC_FFTFILT_POLY_TUNE=$2 timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>>gpurun_out/p_err.txt | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('roofline', {}).get('frac'))
" | tee -a gpurun_out/p_c5_variants.txt
done
for v in "4 0" "4 2"; do
  set -- $v
  echo "== trace C=$1 tune=$2" >> gpurun_out/p_trace.txt
  RRC_FFTFILT_TRACE=1 RRC_FFTFILT_POLY_C=$1 RRC_FFTFILT_POLY_TUNE=$2 timeout 300 python bench.py --config c5 --steps 2 --warmup 1 --headline-only --no-e2e --no-cpu --sustain 0 2>&1 >/dev/null | grep -A15 "iter [456]:" >> gpurun_out/p_trace.txt
done
tail -3 gpurun_out/p_err.txt
