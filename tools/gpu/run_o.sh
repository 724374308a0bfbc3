#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu/tc5_check.py > gpurun_out/o_tc5_check.txt 2>&1; echo "check rc=$?"; tail -60 gpurun_out/o_tc5_check.txt
RRC_FIR_TCGEN05=1 timeout 300 python bench.py --config c1 --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/o_c1_tc5.json 2> gpurun_out/o_c1_tc5.err; echo "bench rc=$?"; cat gpurun_out/o_c1_tc5.json | cut -c1-600; tail -3 gpurun_out/o_c1_tc5.err
