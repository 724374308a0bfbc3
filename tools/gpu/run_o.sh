#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --config c1 --steps 20 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0"
RRC_FIR_TC5_TRACE=1 timeout 300 $B > /dev/null 2> gpurun_out/o_c1_tc5_trace.txt; grep "tc5 \|^   " gpurun_out/o_c1_tc5_trace.txt | head -60 > gpurun_out/r02_c1_tcgen05_trace.txt
timeout 300 $B > gpurun_out/r02_c1_tcgen05_bench.json 2> gpurun_out/o_c1_tc5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc5_kernel -s 3 -c 1 -f -o /tmp/o_c1 $B > gpurun_out/o_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/tools_ncu_summary.py /tmp/o_c1.ncu-rep > gpurun_out/r02_c1_tcgen05_ncu_summary.txt 2>&1
ncu -i /tmp/o_c1.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); d=dict(zip(rows[0],rows[-1]))
for k in ['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','smsp__mem_tensor_reads_op_ldt.sum','smsp__sass_inst_executed_op_tmem_ldt.sum','smsp__sass_inst_executed_op_tmem_stt.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']: print(f'{k:80s} {d.get(k)}')
" >> gpurun_out/r02_c1_tcgen05_ncu_summary.txt
ncu -i /tmp/o_c1.ncu-rep --page source --csv > /tmp/o_src.csv 2>/dev/null; python profiles/tools_sass_hot.py /tmp/o_src.csv 1.0 >> gpurun_out/r02_c1_tcgen05_ncu_summary.txt 2>&1
tail -30 gpurun_out/r02_c1_tcgen05_ncu_summary.txt | cut -c1-200
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/o_pytest_full.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
