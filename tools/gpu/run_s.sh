#!/bin/bash
# GPU call S: polyphase kernel — L1::no_allocate on the gather, with and without resident spectrum rows
mkdir -p gpurun_out
rm -f gpurun_out/s_c5_variants.txt gpurun_out/s_err.txt
for v in "0 0" "0 32" "1 32" "1 0"; do
  set -- $v
  echo "== HRES=$1 TUNE=$2" | tee -a gpurun_out/s_c5_variants.txt
  RRC_FFTFILT_POLY_HRES=$1 RRC_FFTFILT_POLY_TUNE=$2 timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>>gpurun_out/s_err.txt | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('roofline', {}).get('frac'), d.get('gpu_launches'))
" | tee -a gpurun_out/s_c5_variants.txt
done
tail -3 gpurun_out/s_err.txt
