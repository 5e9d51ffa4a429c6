#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu (ST)"; timeout 900 python -m pytest tests/test_st_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 600 python tools/sweep_st.py ${SWEEP_ARGS} > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
