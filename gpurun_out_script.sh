set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "fir or demod or rtl_fm" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_fir.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
for c in c1 c3; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_${c}_v5.json; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_poly -s 2 -c 1 -o gpurun_out/prof_fir_c3_v5 -f python bench.py --config c3 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu3.log 2>&1
