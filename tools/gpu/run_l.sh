#!/bin/bash
# GPU call L: full suite + smoke + ncu of the new C3 kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/l_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/l_suite.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/l_smoke.log 2>&1; tail -1 gpurun_out/l_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_rtu_kernel -s 2 -c 1 -f -o gpurun_out/l_c3_rtu \
   python bench.py --config c3 --steps 2 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/l_ncu.log 2>&1
echo "ncu rc=$?"
