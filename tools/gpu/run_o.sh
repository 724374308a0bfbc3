#!/bin/bash
mkdir -p gpurun_out
RRC_FIR_TCGEN05=2 timeout 300 python tools/gpu/tc5_check.py 2>&1 | tail -3 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05 or config1_full" -x 2>&1 | tail -3 | tee gpurun_out/o_pytest.txt
B="python bench.py --config c1 --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0"
RRC_FIR_TC5_TRACE=1 timeout 300 $B > /dev/null 2> gpurun_out/o_c1_tc5_trace.txt; grep "tc5 kernel" gpurun_out/o_c1_tc5_trace.txt | head -4
for n in 4194304 8388608 16777216 33554432 67108864; do
for v in 1 0; do
RRC_FIR_TCGEN05=$v timeout 300 python bench.py --config c1 --n $n --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/o_tmp.json 2> gpurun_out/o_c1_tc5.err; python -c "import json;d=json.load(open('gpurun_out/o_tmp.json'));print('n $n RRC_FIR_TCGEN05=$v', round(d['ms_per_step']*1000,2),'us', round(d['roofline']['frac'],3), d['roofline']['kernel'][:16])"
done; done 2>&1 | tee gpurun_out/o_c1_tc5_sizes.txt
