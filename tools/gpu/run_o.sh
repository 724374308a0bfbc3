#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=fir_tc5 python tools/gpu/tc5_sanitize.py > gpurun_out/o_sanitize_$tool.txt 2>&1; echo "$tool rc=$?"; grep -i "error summary\|hazard\|Invalid\|done\|========= [A-Z]" gpurun_out/o_sanitize_$tool.txt | sort | uniq -c | head -12
done
for i in 1 2 3 4 5; do timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05 or config1_full" -x 2>&1 | tail -1; done
