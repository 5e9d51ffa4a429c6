#!/bin/bash
# ncu --set full of the marching forward on C5 (cfg 7) and C2 (cfg 6)
bash tools/gpu_ncu.sh c5 r2k_c5 7
bash tools/gpu_ncu.sh c2 r2k_c2 6
