mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ingest.py tests/test_blocks_gpu.py -m gpu -x -q -k "fir or Fir" -s ) > gpurun_out/pytest_fir_tc.log 2>&1; tail -3 gpurun_out/pytest_fir_tc.log
grep "fir_tcc.*tensor" gpurun_out/pytest_fir_tc.log | sort -k6 -g | tail -2
timeout 600 python tools/fir_sweep.py --ctaps 2>&1 | tail -20
