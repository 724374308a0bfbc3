mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fir" -s ) > gpurun_out/pytest_fir_tc.log 2>&1; tail -5 gpurun_out/pytest_fir_tc.log
grep "fir_tc1" gpurun_out/pytest_fir_tc.log | sort -k5 -g | tail -4
for cfg in c1; do
  for v in "RRC_FIR_TENSOR=0" "RRC_FIR_TENSOR=1" "RRC_FIR_TC1=0"; do
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_tc1_kernel -s 3 -c 1 -f -o gpurun_out/c1_tc_v5 python bench.py --config c1 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_c1_tc.log 2>&1
