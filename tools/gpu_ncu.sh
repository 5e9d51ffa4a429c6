#!/bin/bash
# ncu --set full capture of the ST kernels on one workload:  bash tools/gpu_ncu.sh <workload> <out-name> [fwd_cfg] [bwd_cfg]
WL=${1:-c5}; OUT=${2:-prof}
export SRST_ST_FWD_CFG=${3:--1} SRST_ST_BWD_CFG=${4:--1}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 4 --warmup 3 --workload $WL --no-extra --no-cpu > gpurun_out/plain_$OUT.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:st_ -s 6 -c 2 -o gpurun_out/$OUT python bench.py --steps 4 --warmup 3 --workload $WL --no-extra --no-cpu > gpurun_out/ncu_$OUT.log 2>&1
echo "ncu $OUT rc=$?"; tail -2 gpurun_out/ncu_$OUT.log
