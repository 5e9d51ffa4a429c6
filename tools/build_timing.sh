#!/bin/bash
# Tools-only build of the library with phase time stamps compiled in (-DSRST_TIMING): srgan_st_b200/libsrst_timing.so
cd "$(dirname "$0")/../srgan_st_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DSRST_TIMING \
  --shared -Xcompiler -fPIC -o ../libsrst_timing.so srst_cabi.cu && echo "built libsrst_timing.so"
