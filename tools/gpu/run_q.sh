#!/bin/bash
# GPU call Q: per-phase cycle traces of the polyphase kernel variants on config 5
mkdir -p gpurun_out
: > gpurun_out/q_trace.txt
for v in "4 0" "2 0" "1 0" "4 1" "2 1"; do
  set -- $v
  echo "== C=$1 WS=$2" >> gpurun_out/q_trace.txt
  RRC_FFTFILT_TRACE=1 RRC_FFTFILT_POLY_C=$1 RRC_FFTFILT_POLY_WS=$2 timeout 300 python bench.py --config c5 --steps 2 --warmup 1 --headline-only --no-e2e --no-cpu --sustain 0 2>&1 >/dev/null | grep -A15 "iter [45]:" >> gpurun_out/q_trace.txt
done
wc -l gpurun_out/q_trace.txt
