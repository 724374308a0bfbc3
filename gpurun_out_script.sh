mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ingest.py tests/test_blocks_gpu.py -m gpu -x -q -k "fir or Fir" ) > gpurun_out/pytest_fir_tc.log 2>&1; tail -2 gpurun_out/pytest_fir_tc.log
for cfg in c3 c3 c3u8; do timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_x.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1]); print('$cfg', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"; done
env RRC_FIR_TENSOR=0 timeout 300 python bench.py --config c1 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_x.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1]); print('c1 fp32', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
