#!/bin/bash
mkdir -p gpurun_out
RRC_FIR_TCGEN05=3 timeout 300 python tools/gpu/tc5_check.py 2>&1 | tail -3 | cut -c1-400 | tee gpurun_out/o_tc5_check.txt
B="python bench.py --config c1 --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0"
RRC_FIR_TCGEN05=3 RRC_FIR_TC5_TRACE=1 timeout 300 $B > /dev/null 2> gpurun_out/o_c1_tc5t_trace.txt; grep "tc5p" gpurun_out/o_c1_tc5t_trace.txt | head -18
for v in 3 0 3 0; do
RRC_FIR_TCGEN05=$v timeout 300 $B > gpurun_out/o_c1_tc5_$v.json 2> gpurun_out/o_c1_tc5.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/o_c1_tc5_$v.json'));print('variant $v', d['ms_per_step'],d['roofline']['frac'], d['roofline']['kernel'][:20])"
done
RRC_FIR_TCGEN05=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc5t_kernel -s 3 -c 1 -f -o /tmp/o_c1 $B > gpurun_out/o_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/tools_ncu_summary.py /tmp/o_c1.ncu-rep > gpurun_out/o_c1_tc5t_ncu_summary.txt 2>&1; cat gpurun_out/o_c1_tc5t_ncu_summary.txt
ncu -i /tmp/o_c1.ncu-rep --page source --csv > /tmp/o_src.csv 2>/dev/null; python profiles/tools_sass_hot.py /tmp/o_src.csv 0.8 > gpurun_out/o_c1_tc5t_sass_hot.txt 2>&1; cat gpurun_out/o_c1_tc5t_sass_hot.txt | cut -c1-220
