mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ingest.py tests/test_blocks_gpu.py -m gpu -x -q -k "fir or Fir" -s ) > gpurun_out/pytest_fir_tc.log 2>&1; tail -2 gpurun_out/pytest_fir_tc.log
grep "fir_tc1" gpurun_out/pytest_fir_tc.log | sort -k5 -g | tail -2
timeout 600 python tools/fir_sweep.py 2>&1 | tail -20
for v in "A=1" "A=2"; do
    timeout 300 python bench.py --config c1 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c1_occ.json 2> gpurun_out/bench_c1_occ.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c1_occ.json').read().strip().splitlines()[-1]); print('c1', '$v', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))
PY
done
