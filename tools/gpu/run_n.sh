#!/bin/bash
# GPU call N: default bench (all configs) + ncu launch list of the same command + ncu --set full per headline kernel,
# summarised on the box (the reports together exceed the 64 MiB that travel back)
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -3 gpurun_out/n_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/n_launches.log 2>&1; echo "launch list rc=$?"
for spec in "c1:fir_tc1_kernel" "c2:fftfilt_tma_kernel" "c3:fir_rtu_kernel" "c4:resample_kernel" "c5:fftfilt_fold_kernel"; do
  c=${spec%%:*}; k=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/n_$c \
     python bench.py --config $c --steps 2 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/n_ncu_$c.log 2>&1; echo "ncu $c rc=$?"
  python profiles/tools_ncu_summary.py /tmp/n_$c.ncu-rep > gpurun_out/n_${c}_ncu_summary.txt 2>&1
done
