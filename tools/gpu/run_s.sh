#!/bin/bash
# GPU call S: polyphase kernel — second launch on the idle SMs (split weight sweep), then parity at size
mkdir -p gpurun_out
rm -f gpurun_out/s_c5_variants.txt gpurun_out/s_err.txt
for v in "1 50" "1 58" "1 64"; do
  set -- $v
  echo "== SPLIT=$1 W=$2" | tee -a gpurun_out/s_c5_variants.txt
  RRC_FFTFILT_POLY_SPLIT=$1 RRC_FFTFILT_POLY_SPLIT_W=$2 timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>>gpurun_out/s_err.txt | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('roofline', {}).get('frac'), d.get('gpu_launches'))
" | tee -a gpurun_out/s_c5_variants.txt
done
tail -3 gpurun_out/s_err.txt
