#!/bin/bash
# GPU call R: ncu --set full of the polyphase kernel on config 5
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fftfilt_poly_kernel -s 3 -c 1 -f -o /tmp/r_c5 \
   python bench.py --config c5 --steps 2 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/r_ncu_c5.log 2>&1; echo "ncu rc=$?"
python profiles/tools_ncu_summary.py /tmp/r_c5.ncu-rep > gpurun_out/r_c5_ncu_summary.txt 2>&1
ncu -i /tmp/r_c5.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin))
hdr, vals = rows[0], rows[-1]
for h, v in zip(hdr, vals):
    if any(k in h for k in ('dram__bytes', 'lts__t_sector', 'lts__t_bytes', 'hit_rate', 'l1tex__t_sectors_pipe_lsu_mem_global', 'lts__t_sectors_srcunit_tex', 'lts__d_sectors_fill')):
        print(h, v)
" > gpurun_out/r_c5_raw_mem.txt
wc -l gpurun_out/r_c5_raw_mem.txt
