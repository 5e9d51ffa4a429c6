#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2f_*.ncu-rep
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== host overhead"; python tools/host_overhead.py 2>&1 | tee gpurun_out/host_overhead.log | head -3
python tools/bench_stpx.py 2>&1 | tee gpurun_out/stpx.log
bash tools/gpu_ncu.sh c2 r2f_c2
bash tools/gpu_ncu.sh c5 r2f_c5
bash tools/gpu_ncu.sh c2x r2f_c2x
