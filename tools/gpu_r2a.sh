#!/bin/bash
# Round-2 first GPU pass: ST parity tests on the new kernels, then the tile sweep.
mkdir -p gpurun_out
echo "== pytest ST"; timeout 1200 python -m pytest tests/test_st_gpu.py tests/test_stpx.py -m gpu -q -x > gpurun_out/pytest_st.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_st.log
echo "== sweep"; timeout 900 python tools/sweep_st.py > gpurun_out/sweep_r2a.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_r2a.log
