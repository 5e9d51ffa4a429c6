#!/bin/bash
mkdir -p gpurun_out
echo "== pytest ST"; timeout 1200 python -m pytest tests/test_st_gpu.py tests/test_stpx.py -m gpu -q -x > gpurun_out/pytest_st.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_st.log
echo "== sweep bwd"; SWEEP_FWD=1 SWEEP_SHAPES=64x96x96,1024x96x96,1x1356x2040,4x1356x2040 timeout 900 python tools/sweep_st.py 2>&1 | grep "bwd cfg\|default"
