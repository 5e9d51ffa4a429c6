#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
bash tools/gpu_ncu.sh c2 r2f_c2
bash tools/gpu_ncu.sh c5 r2f_c5
