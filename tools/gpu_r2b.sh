#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2b_*.ncu-rep
bash tools/gpu_ncu.sh c2 r2b_c2_f0 0 1
bash tools/gpu_ncu.sh c2 r2b_c2_f2 2 2
bash tools/gpu_ncu.sh c5 r2b_c5_f1 1 0
echo "== sweep noixy"; SWEEP_NOIXY=1 SWEEP_BWD=0 SWEEP_SHAPES=64x96x96,1x1356x2040 timeout 600 python tools/sweep_st.py 2>&1 | grep fwd
