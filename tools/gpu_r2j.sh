#!/bin/bash
# round 2, marching forward: parity first, then the forward sweep (tiled cfgs 1,3,4 vs marching 6,7)
mkdir -p gpurun_out
echo "== pytest st gpu"; timeout 900 python -m pytest tests/test_st_gpu.py tests/test_stpx.py -m gpu -q -x > gpurun_out/pytest_st.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_st.log
echo "== sweep"; SWEEP_FWD=1,3,4,6,7 SWEEP_BWD=6,7 timeout 600 python tools/sweep_st.py > gpurun_out/sweep_march.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep_march.log
