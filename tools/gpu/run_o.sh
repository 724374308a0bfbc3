#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05" -x 2>&1 | tail -5 | tee gpurun_out/o_pytest.txt
B="python bench.py --config c1 --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0"
for c in 1 2; do
RRC_FIR_TC5_CTAS=$c RRC_FIR_TCGEN05=1 RRC_FIR_TC5_TRACE=1 timeout 300 $B > /dev/null 2> gpurun_out/o_c1_tc5_trace$c.txt; grep "tc5:" gpurun_out/o_c1_tc5_trace$c.txt | head -3; grep -A8 "tc5 trace" gpurun_out/o_c1_tc5_trace$c.txt | head -18
RRC_FIR_TC5_CTAS=$c RRC_FIR_TCGEN05=1 timeout 300 $B > gpurun_out/o_c1_tc5_$c.json 2> gpurun_out/o_c1_tc5.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/o_c1_tc5_$c.json'));print('CTAS $c', d['ms_per_step'],d['roofline'])"
done
RRC_FIR_TC5_CTAS=2 RRC_FIR_TCGEN05=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc5_kernel -s 3 -c 1 -f -o /tmp/o_c1 $B > gpurun_out/o_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/tools_ncu_summary.py /tmp/o_c1.ncu-rep > gpurun_out/o_c1_tc5_ncu_summary.txt 2>&1; cat gpurun_out/o_c1_tc5_ncu_summary.txt
