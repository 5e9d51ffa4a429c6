#!/bin/bash
mkdir -p gpurun_out
echo "== pytest live ref + st"; timeout 1500 python -m pytest tests/test_ref_live_gpu.py tests/test_st_gpu.py -m gpu -q > gpurun_out/pytest_ref.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ref.log
echo "== sweep"; SWEEP_BWD=7 SWEEP_SHAPES=64x96x96,1024x96x96,1x1356x2040,4x1356x2040 timeout 900 python tools/sweep_st.py 2>&1 | grep "fwd cfg\|default" | tee gpurun_out/sweep_r2h.log
