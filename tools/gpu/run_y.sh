#!/bin/bash
# GPU call Y: final tree (FftFilter variant 42 default) — full GPU suite, smoke, default bench, ncu launch list, ncu --set full of config 2
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/y_pytest_full.txt 2>&1; grep -E "passed|failed" gpurun_out/y_pytest_full.txt | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/y_smoke.txt
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err; tail -3 gpurun_out/y_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/y_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 --no-e2e > gpurun_out/y_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftfilt_tmh_kernel -s 3 -c 1 -f -o /tmp/y_c2 \
   python bench.py --config c2 --steps 2 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/y_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
python profiles/tools_ncu_summary.py /tmp/y_c2.ncu-rep > gpurun_out/y_c2_ncu_summary.txt 2>&1
