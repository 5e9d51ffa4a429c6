#!/bin/bash
# live-reference parity (oracle/_ref on the B200) + ncu --set full of the default kernels on C2 and C5
mkdir -p gpurun_out
echo "== pytest live ref"; timeout 900 python -m pytest tests/test_ref_live_gpu.py -m gpu -q > gpurun_out/pytest_live.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_live.log
bash tools/gpu_ncu.sh c5 r2m_c5
bash tools/gpu_ncu.sh c2 r2m_c2
