#!/bin/bash
# GPU call T: round-2 final state — full GPU test suite, default bench (all configs) + reference arm, ncu launch list of the
# default command, ncu --set full of the config-5 polyphase kernel, per-phase trace
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/t_pytest_full.txt 2>&1; tail -3 gpurun_out/t_pytest_full.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/t_smoke.txt
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; tail -3 gpurun_out/t_bench.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/t_bench_reference.json 2> gpurun_out/t_bench_reference.err; tail -3 gpurun_out/t_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/t_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftfilt_poly_kernel -s 3 -c 1 -f -o /tmp/t_c5 \
   python bench.py --config c5 --steps 2 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/t_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
python profiles/tools_ncu_summary.py /tmp/t_c5.ncu-rep > gpurun_out/t_c5_ncu_summary.txt 2>&1
RRC_FFTFILT_TRACE=1 timeout 300 python bench.py --config c5 --steps 2 --warmup 1 --headline-only --no-e2e --no-cpu --sustain 0 2>&1 >/dev/null | grep -A15 "iter [4567]:" > gpurun_out/t_c5_trace.txt
