#!/bin/bash
mkdir -p gpurun_out
echo "== pytest live ref"; timeout 1500 python -m pytest tests/test_ref_live_gpu.py -m gpu -q > gpurun_out/pytest_ref.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ref.log
echo "== phase timing"
python tools/phase_timing.py 64 96 96 2
python tools/phase_timing.py 64 96 96 3
python tools/phase_timing.py 1024 96 96 6
python tools/phase_timing.py 1 1356 2040 1
