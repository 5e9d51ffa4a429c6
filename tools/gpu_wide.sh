#!/bin/bash
mkdir -p gpurun_out
echo "== pytest ST gpu"; timeout 900 python -m pytest tests/test_st_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== sweep PDL=1"; SWEEP_STREAM=0 SWEEP_FWD=${SWEEP_FWD:-2,10} SWEEP_BWD_MAX=${SWEEP_BWD_MAX:-1} timeout 600 python tools/sweep_st.py > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
echo "== sweep PDL=0"; SRST_PDL=0 SWEEP_STREAM=0 SWEEP_FWD=${SWEEP_FWD:-2,10} SWEEP_BWD_MAX=${SWEEP_BWD_MAX:-1} timeout 600 python tools/sweep_st.py > gpurun_out/sweep_nopdl.log 2>&1; cat gpurun_out/sweep_nopdl.log
echo "== stamps"; timeout 200 python tools/wide_debug.py 2>&1 | head -16 | tee gpurun_out/wide_debug.log
