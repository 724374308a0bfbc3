#!/bin/bash
# GPU call S: polyphase kernel experiments on config 5
mkdir -p gpurun_out
rm -f gpurun_out/s_c5_variants.txt gpurun_out/s_err.txt
for v in "0 0" "1 0"; do
  set -- $v
  echo "== C=4 HRES=$1 PREH=$2" | tee -a gpurun_out/s_c5_variants.txt
  RRC_FFTFILT_POLY_HRES=$1 RRC_FFTFILT_POLY_PREH=$2 timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>>gpurun_out/s_err.txt | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('roofline', {}).get('frac'))
" | tee -a gpurun_out/s_c5_variants.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "polyphase" 2>&1 | tail -3 | tee gpurun_out/s_pytest.txt
RRC_FFTFILT_POLY_PREH=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "polyphase_decimation and 16385" 2>&1 | tail -3 | tee -a gpurun_out/s_pytest.txt
tail -3 gpurun_out/s_err.txt
