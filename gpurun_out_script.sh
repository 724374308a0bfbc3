mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python tools/fir_sweep.py > /dev/null 2>&1; timeout 600 python tools/fir_sweep.py --f32 > /dev/null 2>&1; timeout 600 python tools/fir_sweep.py --ctaps 2>&1 | head -3
