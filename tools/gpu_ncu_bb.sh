#!/bin/bash
# ncu --set full capture of the Best-Buddy search kernel on BASELINE config 4 (B=64, 192x192).
mkdir -p gpurun_out
timeout 600 python tools/bench_bb.py > gpurun_out/plain_bb.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bb_search -s 3 -c 1 -o gpurun_out/r2n_bb python tools/bench_bb.py > gpurun_out/ncu_bb.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_bb.log; cat gpurun_out/plain_bb.log
