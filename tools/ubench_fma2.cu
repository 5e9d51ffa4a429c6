// Micro-benchmark of the exact FFMA2 operand forms the ST kernels use.
//   A: acc2 = v2 * bcast(k_const) + acc2          (horizontal passes: R.F32x2, UR.F32, R.F32x2)
//   B: acc2 = bcast(v) * kpair_const + acc2       (vertical passes:   R.F32,  UR.F32x2, R.F32x2)
//   C: form A with an independent LDS.128 every 4 FFMA2 (phase-D-like mix)
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
struct Taps { float k[16]; float2 kp[16]; };
constexpr int ITERS = 1024, NACC = 16;

template <int FORM>
__global__ void __launch_bounds__(256) kern(float* out, long long* cyc, const __grid_constant__ Taps t, float x0) {
  __shared__ float4 sm[512];
  sm[threadIdx.x] = make_float4(x0, x0 * 2, x0 * 3, x0 * 4);
  sm[threadIdx.x + 256] = make_float4(x0, x0, 1.f, 0.f);
  __syncthreads();
  float2 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  float2 v = make_float2(x0 + threadIdx.x * 1e-6f, x0);
  float4 w = sm[threadIdx.x];
  int idx = threadIdx.x;
  long long c0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (FORM == 0) acc[i] = __ffma2_rn(v, make_float2(t.k[i], t.k[i]), acc[i]);
      if (FORM == 1) acc[i] = __ffma2_rn(make_float2(v.x, v.x), t.kp[i], acc[i]);
      if (FORM == 2) {
        acc[i] = __ffma2_rn(make_float2(w.x, w.y), make_float2(t.k[i], t.k[i]), acc[i]);
        if ((i & 3) == 3) { idx = (idx + 32) & 511; w = sm[idx]; }   // independent of acc: only the NEXT group uses it
      }
      if (FORM == 3) {
        acc[i] = __ffma2_rn(make_float2(w.x, w.x), t.kp[i], acc[i]);
        if ((i & 7) == 7) { idx = (idx + 32) & 511; float2 u = *reinterpret_cast<float2*>(&sm[idx]); w.x = u.x; w.y = u.y; }
      }
    }
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + w.z;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

int main() {
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(float) * 256 * 8 * nsm)); CK(cudaMalloc(&d_cyc, sizeof(long long) * 8 * nsm));
  Taps t; for (int i = 0; i < 16; ++i) { t.k[i] = 1e-3f * (i + 1); t.kp[i] = make_float2(1e-3f * i, 2e-3f * i); }
  const char* names[4] = {"A: pair * bcast(const)      ", "B: bcast(reg) * const pair  ", "C: form A + LDS.128 / 4 FFMA2", "D: form B + LDS.64 / 8 FFMA2 "};
  for (int bps : {1, 2, 3}) for (int f = 0; f < 4; ++f) {
    int nb = bps * nsm;
    auto launch = [&]() { if (f == 0) kern<0><<<nb, 256>>>(d_out, d_cyc, t, 0.5f); if (f == 1) kern<1><<<nb, 256>>>(d_out, d_cyc, t, 0.5f);
                          if (f == 2) kern<2><<<nb, 256>>>(d_out, d_cyc, t, 0.5f); if (f == 3) kern<3><<<nb, 256>>>(d_out, d_cyc, t, 0.5f); };
    launch(); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane = 2.0 * ITERS * NACC * 256.0 * nb;
    printf("%s blocks/SM=%d  %.3f ms  %.2f T lane-FMA/s  (%.1f lane-FMA/clk/SM at 1.965 GHz)\n", names[f], bps, ms, lane / (ms * 1e-3) / 1e12, lane / (ms * 1e-3) / 1.965e9 / nsm);
  }
  return 0;
}
