#!/bin/bash
# GPU call X: FftFilter variant 40 (spectrum in tensor memory, now the default) — full GPU suite, config 2 against variant 36
mkdir -p gpurun_out
rm -f gpurun_out/x_variants.txt gpurun_out/x_err.txt
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/x_pytest_full.txt 2>&1; grep -E "passed|failed" gpurun_out/x_pytest_full.txt | tail -2
for v in 36 40 36 40; do
  echo "== variant $v" | tee -a gpurun_out/x_variants.txt
  RRC_FFTFILT_VARIANT=$v timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>>gpurun_out/x_err.txt | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('roofline', {}).get('frac'), d.get('roofline', {}).get('kernel', '')[:40])
" | tee -a gpurun_out/x_variants.txt
done
tail -3 gpurun_out/x_err.txt
