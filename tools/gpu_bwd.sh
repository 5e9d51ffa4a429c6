#!/bin/bash
mkdir -p gpurun_out
echo "== ST gpu tests (TMA backward)"; timeout 600 python -m pytest tests/test_st_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "== sweep tma=1"; SWEEP_STREAM=0 SWEEP_FWD=0,1,2 SWEEP_BWD_MAX=2 timeout 300 python tools/sweep_st.py 2>&1 | grep -E "B= *(64|256|1|4) "
echo "== sweep tma=0"; SRST_ST_BWD_TMA=0 SWEEP_STREAM=0 SWEEP_FWD=0,1,2 SWEEP_BWD_MAX=2 timeout 300 python tools/sweep_st.py 2>&1 | grep -E "B= *(64|256|1|4) "
