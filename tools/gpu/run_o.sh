#!/bin/bash
mkdir -p gpurun_out
RRC_FIR_TCGEN05=3 timeout 300 python tools/gpu/tc5_check.py 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/o_tc5_check.txt
B="python bench.py --config c1 --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0"
RRC_FIR_TCGEN05=3 RRC_FIR_TC5_TRACE=1 timeout 300 $B > /dev/null 2> gpurun_out/o_c1_tc5t_trace.txt; grep "tc5p kernel\|tc5p prologue" gpurun_out/o_c1_tc5t_trace.txt | head -40
for v in 3 0 3 0; do
RRC_FIR_TCGEN05=$v timeout 300 $B > gpurun_out/o_c1_tc5_$v.json 2> gpurun_out/o_c1_tc5.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/o_c1_tc5_$v.json'));print('variant $v', d['ms_per_step'],d['roofline']['frac'], d['roofline']['kernel'][:20])"
done
