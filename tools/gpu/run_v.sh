#!/bin/bash
# GPU call V: compute-sanitizer on the polyphase kernel
mkdir -p gpurun_out
timeout 200 python tools/gpu/poly_sanitize.py 2>&1 | tail -12 | tee gpurun_out/v_plain.txt
for tool in memcheck synccheck racecheck; do
  POLY_SANITIZE_SMALL=1 timeout 1200 compute-sanitizer --tool $tool --kernel-regex kns=fftfilt_poly python tools/gpu/poly_sanitize.py > gpurun_out/v_sanitize_$tool.txt 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/v_sanitize_$tool.txt
done
