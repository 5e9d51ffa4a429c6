#!/bin/bash
# ncu --set full of the stand-alone marching-forward driver:  bash tools/gpu_ncu_march.sh <tw> <B> <H> <W> <out>
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 50 -c 1 -o gpurun_out/$5 tools/_bin/march_bench_$1 $2 $3 $4 0 0 > gpurun_out/ncu_$5.log 2>&1
echo "ncu $5 rc=$?"; tail -2 gpurun_out/ncu_$5.log
