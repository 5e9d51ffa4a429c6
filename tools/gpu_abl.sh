#!/bin/bash
# phase ablation of the marching forward (tools/march_bench.cu): what does each phase cost?
mkdir -p gpurun_out
M="${MASKS:-0 1 3 28 127}"
{ for v in ${VARS:-112 112_4 96_4 80_4 64_4}; do tools/_bin/march_bench_$v 1 1356 2040 0 $M; done
  for v in ${VARS2:-96 96_4}; do tools/_bin/march_bench_$v 64 96 96 0 $M; tools/_bin/march_bench_$v 1024 96 96 0 0 1 3 28 127; done; } 2>&1 | tee gpurun_out/march_abl.log
