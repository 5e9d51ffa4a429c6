#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bb bench"; timeout 300 python tools/bench_bb.py 2>&1 | tee gpurun_out/bench_bb.log
