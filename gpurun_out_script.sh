mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
grep -n "^E " gpurun_out/pytest_gpu_all.log | head -5
timeout 600 python tools/fir_sweep.py 2>&1 | grep "D=  8"
timeout 600 python tools/fir_sweep.py --f32 2>&1 | grep "D=  8"
timeout 600 python tools/fir_sweep.py --ctaps 2>&1 | grep "D=  8"
