mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ingest.py -m gpu -x -q 2>&1 | tail -30
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
