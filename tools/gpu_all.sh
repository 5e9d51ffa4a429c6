#!/bin/bash
# All GPU tests, then the ST tile sweep of the library's default configurations.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== sweep"; SWEEP_STREAM=0 SWEEP_FWD=${SWEEP_FWD:-2,3} SWEEP_BWD_MAX=${SWEEP_BWD_MAX:-3} timeout 600 python tools/sweep_st.py > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
