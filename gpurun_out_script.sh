mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fir" -s ) > gpurun_out/pytest_fir_tc.log 2>&1; tail -3 gpurun_out/pytest_fir_tc.log
grep "fir_tc1" gpurun_out/pytest_fir_tc.log | sort -k5 -g | tail -3
for cfg in c1; do
  for v in "RRC_FIR_TENSOR=1"; do
    tag=$(echo "$v" | tr ' =' '__')
    env $v timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_${tag}.json 2> gpurun_out/bench_${cfg}_${tag}.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${cfg}_${tag}.json').read().strip().splitlines()[-1]); print('$cfg', '$v', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), d['roofline']['kernel'][:30])
except Exception as e: print('$cfg $v failed', e)
PY
  done
done
