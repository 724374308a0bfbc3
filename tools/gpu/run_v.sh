#!/bin/bash
# GPU call V: compute-sanitizer on the TMEM-table FftFilter kernel
mkdir -p gpurun_out
timeout 200 python tools/gpu/tmh_sanitize.py 2>&1 | tail -14 | tee gpurun_out/v_tmh_plain.txt
for tool in memcheck synccheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --kernel-regex kns=fftfilt_tmh python tools/gpu/tmh_sanitize.py > gpurun_out/v_tmh_sanitize_$tool.txt 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/v_tmh_sanitize_$tool.txt
done
