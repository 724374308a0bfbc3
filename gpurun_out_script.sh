mkdir -p gpurun_out
timeout 300 python bench.py --config c1 --steps 20 --warmup 3 > gpurun_out/bench_c1_v10.json 2> gpurun_out/bench_c1_v10.err; tail -c 600 gpurun_out/bench_c1_v10.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_tc1_kernel -s 3 -c 1 -f -o gpurun_out/c1_tc_v6 python bench.py --config c1 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_c1_tc.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_c1_v10.csv python bench.py --config c1 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
