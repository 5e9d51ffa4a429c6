#!/bin/bash
# Phase ablation and phase stamps of the marching forward (tools/march_bench.cu): what does each phase cost, where does a block's time go?
# mask bits (set = skipped): 1 chain, 2 vertical, 4 horizontal, 8 gradient, 16 gray conversion, 32 ds stores, 64 Ix/Iy stores
mkdir -p gpurun_out
M="${MASKS:-0 1 2 3 4 8 16 28 32 64 96 127}"
{ tools/_bin/march_bench_112_8 1 1356 2040 0 $M; tools/_bin/march_bench_96_8 64 96 96 0 $M; tools/_bin/march_bench_96_8 1024 96 96 0 0 1 3 28 127; } 2>&1 | tee gpurun_out/march_abl.log
MB_STAMPS=1 tools/_bin/march_stamps_96_8 64 96 96 0 0 > gpurun_out/march_stamps_c2.log 2>&1
MB_STAMPS=1 tools/_bin/march_stamps_112_8 1 1356 2040 0 0 > gpurun_out/march_stamps_c5.log 2>&1
