#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/m_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/m_suite.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/m_smoke.log 2>&1; tail -1 gpurun_out/m_smoke.log
