#!/bin/bash
# A/B the working-tree library against srgan_st_b200/libsrst_ab.so on one box, then the ST GPU tests.
for i in 1 2; do
  echo NEW; SWEEP_STREAM=0 SWEEP_FWD=${SWEEP_FWD:-2,3} SWEEP_BWD_MAX=3 python tools/sweep_st.py 2>&1 | grep -v "B= 16\|NVIDIA"
  echo OLD; SRST_LIB=srgan_st_b200/libsrst_ab.so SWEEP_STREAM=0 SWEEP_FWD=${SWEEP_FWD:-2,3} SWEEP_BWD_MAX=3 python tools/sweep_st.py 2>&1 | grep -v "B= 16\|NVIDIA"
done
python -m pytest tests/test_st_gpu.py -m gpu -q -x 2>&1 | tail -2
