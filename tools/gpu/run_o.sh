#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05 or config1_full or f32_streams" -x 2>&1 | tail -4 | tee gpurun_out/o_pytest.txt
for n in 8388608 33554432 134217728; do
for v in 1 0; do
RRC_FIR_TCGEN05=$v timeout 300 python bench.py --config c1f --n $n --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/o_tmp.json 2> gpurun_out/o_c1_tc5.err; python -c "import json;d=json.load(open('gpurun_out/o_tmp.json'));print('c1f n $n RRC_FIR_TCGEN05=$v', round(d['ms_per_step']*1000,2),'us', round(d['roofline']['frac'],3), d['roofline']['kernel'][:16])" || tail -3 gpurun_out/o_c1_tc5.err
done; done 2>&1 | tee gpurun_out/o_c1f_sizes.txt
