#!/bin/bash
mkdir -p gpurun_out
for n in 16777216 67108864; do
for v in 1 0; do
RRC_FIR_TCGEN05=$v timeout 300 python bench.py --config c1 --n $n --steps 30 --warmup 8 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/o_tmp.json 2> gpurun_out/o_c1_tc5.err; python -c "import json;d=json.load(open('gpurun_out/o_tmp.json'));print('n $n RRC_FIR_TCGEN05=$v', round(d['ms_per_step']*1000,2),'us', round(d['roofline']['frac'],3), d['roofline']['kernel'][:16])"; tail -2 gpurun_out/o_c1_tc5.err
done; done 2>&1 | tee gpurun_out/o_c1_rot.txt
B="python bench.py --config c1 --steps 20 --warmup 8 --headline-only --no-e2e --no-cpu --sustain 0"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fir_tc5_kernel -s 8 -c 6 --csv $B 2>/dev/null | grep -v "^==" | tail -20 | cut -c1-200
