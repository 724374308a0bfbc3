#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_rtu_kernel -s 2 -c 1 -f -o gpurun_out/j_c3_rtu \
   python bench.py --config c3 --steps 2 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/j_ncu.log 2>&1
echo "ncu rc=$?"
