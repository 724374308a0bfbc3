#!/bin/bash
# GPU call I: uniform-tap FIR kernel (fir_rtu_kernel): parity + config 3 timing; c2 re-check.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ingest.py tests/test_blocks_gpu.py -m gpu -x -q -k "fir or Fir or ingest or rtl" > gpurun_out/i_parity.log 2>&1; echo "fir parity rc=$?"; tail -2 gpurun_out/i_parity.log
timeout 900 python -m pytest tests/test_baseline_size.py -m gpu -x -q -s -k "config3" > gpurun_out/i_c3size.log 2>&1; echo "c3 size rc=$?"; grep "config 3 full" gpurun_out/i_c3size.log; tail -1 gpurun_out/i_c3size.log
for cfg in c3 c3u8; do for env in "RRC_FIR_RTU=0" "RRC_FIR_RTU_R=8" "RRC_FIR_RTU_R=8 RRC_FIR_RTU_NT=128" "RRC_FIR_RTU_R=16" "RRC_FIR_RTU_R=16 RRC_FIR_RTU_NT=64"; do
  env $env timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg $env', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['roofline']['fp32']['frac'],3))"
done; done
timeout 300 python bench.py --config c2 --steps 30 --warmup 5 --headline-only --no-e2e --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
