// Common device-side helpers for the srst kernels (sm_100a).
//
// Compiles two ways:
//   nvcc  -gencode arch=compute_100a,code=sm_100a   -> the product (libsrst.so)
//   g++   -DSRST_EMULATE                            -> tests/emu host emulation (test infra only;
//                                                      one OS thread per CUDA thread, used to
//                                                      check index logic without a GPU)
#pragma once

#ifdef SRST_EMULATE
#include "cuda_emu.h"
#define SRST_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::g_block->dyn_smem)
#define SRST_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define SRST_SET_SMEM(kernel, bytes) (0)
#else
#include <cuda_runtime.h>
#define SRST_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; \
  T* name = reinterpret_cast<T*>(name##_raw_)
#define SRST_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define SRST_SET_SMEM(kernel, bytes) \
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif

#define SRST_DEV __device__ __forceinline__

namespace srst {

constexpr int round_up4(int x) { return (x + 3) / 4 * 4; }
constexpr int round_dn4(int x) { return x / 4 * 4; }
// Shared-memory row pitch (floats): >= w, multiple of 4 (16-byte rows for LDS.128) and == 4 mod 8,
// so that 8 lanes reading float4s from 8 consecutive rows hit 8 distinct 16-byte bank groups.
constexpr int smem_pitch(int w) {
  int p = round_up4(w);
  while (p % 8 != 4) p += 4;
  return p;
}
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int cmin(int a, int b) { return a < b ? a : b; }

SRST_DEV float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
SRST_DEV void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
SRST_DEV float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// torchvision rgb_to_grayscale weights (reference loss.py:400-401).
constexpr float kGrayR = 0.2989f, kGrayG = 0.587f, kGrayB = 0.114f;
SRST_DEV float gray_of(float r, float g, float b) { return (kGrayR * r + kGrayG * g) + kGrayB * b; }

SRST_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace srst
