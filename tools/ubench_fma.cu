// Micro-benchmark: FP32 issue rates on B200 that decide the structure-tensor kernel design.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fma tools/ubench_fma.cu && ./ubench_fma
// Reports lane-FMAs per clock per SM for: FFMA with a constant-bank tap, FFMA with three register
// operands, packed FFMA2 (fma.rn.f32x2) with a scalar tap, FFMA2 + LDS.128 mixed, and MUFU ops.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Taps { float k[16]; };
constexpr int ITERS = 2048;
constexpr int NACC = 16;

__global__ void __launch_bounds__(256) k_ffma_const(float* out, long long* cyc, const __grid_constant__ Taps t, float x0) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  float x = x0 + threadIdx.x * 1e-6f;
  long long c0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(x, t.k[i], acc[i]);
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

__global__ void __launch_bounds__(256) k_ffma_reg(float* out, long long* cyc, const float* in, float x0) {
  float acc[NACC], w[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = threadIdx.x * 1e-3f + i; w[i] = in[i]; }
  float x = x0 + threadIdx.x * 1e-6f;
  long long c0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(x, w[i], acc[i]);
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

__global__ void __launch_bounds__(256) k_ffma2_const(float* out, long long* cyc, const __grid_constant__ Taps t, float x0) {
  float2 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  float2 x = make_float2(x0 + threadIdx.x * 1e-6f, x0);
  long long c0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(x, make_float2(t.k[i], t.k[i]), acc[i]);
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

__global__ void __launch_bounds__(256) k_ffma2_reg(float* out, long long* cyc, const float* in, float x0) {
  float2 acc[NACC], w[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i); w[i] = make_float2(in[i], in[i + 1]); }
  float2 x = make_float2(x0 + threadIdx.x * 1e-6f, x0);
  long long c0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(x, w[i], acc[i]);
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

// FFMA (const tap) + one LDS.128 per LDS_EVERY FMAs: does the load steal FMA issue slots?
template <int LDS_EVERY, bool PACKED>
__global__ void __launch_bounds__(256) k_mix(float* out, long long* cyc, const __grid_constant__ Taps t, float x0) {
  __shared__ float4 sm[256 * 2];
  sm[threadIdx.x] = make_float4(x0, x0, x0, x0);
  sm[threadIdx.x + 256] = make_float4(x0, x0, 1.f, 0.f);
  __syncthreads();
  float2 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  float4 v = sm[threadIdx.x];
  int idx = threadIdx.x;
  long long c0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (PACKED) {
        acc[i] = __ffma2_rn(make_float2(v.x, v.y), make_float2(t.k[i], t.k[i]), acc[i]);
      } else {
        acc[i].x = fmaf(v.x, t.k[i], acc[i].x);
        acc[i].y = fmaf(v.y, t.k[i], acc[i].y);
      }
      if (i % LDS_EVERY == LDS_EVERY - 1) {
        idx = (idx + 32) & 511;
        float4 nv = sm[idx];
        v.x += nv.w; v.y = nv.y;
      }
    }
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + v.x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int OP>
__global__ void __launch_bounds__(256) k_mufu(float* out, long long* cyc, float x0) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 1.5f + threadIdx.x * 1e-3f + i;
  long long c0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (OP == 0) acc[i] = rsqrtf(acc[i]);          // MUFU.RSQ (+ fixup)
      if (OP == 1) acc[i] = __log2f(acc[i]) + 3.f;   // MUFU.LG2
      if (OP == 2) acc[i] = logf(acc[i]) + 3.f;      // full-precision logf
      if (OP == 3) acc[i] = sqrtf(acc[i]) + 1.f;     // IEEE sqrt
      if (OP == 4) acc[i] = 1.0f / acc[i] + 1.f;     // IEEE divide
      if (OP == 5) acc[i] = __fdividef(1.0f, acc[i]) + 1.f;
    }
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <class F>
void run(const char* name, double lane_ops_per_thread, int blocks_per_sm, int nsm, F launch, long long* d_cyc) {
  const int nblk = blocks_per_sm * nsm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(nblk);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  launch(nblk);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(nblk);
  CK(cudaMemcpy(cyc.data(), d_cyc, nblk * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0; long long mx = 0;
  for (auto c : cyc) { mean += c; if (c > mx) mx = c; }
  mean /= nblk;
  const double ops_per_sm = lane_ops_per_thread * 256.0 * blocks_per_sm;
  printf("%-34s blocks/SM=%d  time=%.3f ms  cycles(mean)=%.0f  lane-ops/clk/SM=%.1f  chip=%.2f Tops/s\n", name,
         blocks_per_sm, ms, mean, ops_per_sm / mean, lane_ops_per_thread * 256.0 * nblk / (ms * 1e-3) / 1e12);
}

int main() {
  int dev = 0, nsm = 0, clk = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
  printf("SMs=%d  clockRate=%d kHz\n", nsm, clk);
  float* d_out; long long* d_cyc; float* d_in;
  CK(cudaMalloc(&d_out, sizeof(float) * 256 * 8 * nsm));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 8 * nsm));
  CK(cudaMalloc(&d_in, sizeof(float) * 64));
  std::vector<float> h(64);
  for (int i = 0; i < 64; ++i) h[i] = 1e-3f * (i + 1);
  CK(cudaMemcpy(d_in, h.data(), sizeof(float) * 64, cudaMemcpyHostToDevice));
  Taps t;
  for (int i = 0; i < 16; ++i) t.k[i] = 1e-3f * (i + 1);
  const double n1 = (double)ITERS * NACC;
  for (int bps : {1, 2, 4}) {
    run("FFMA  tap=const", n1, bps, nsm, [&](int nb) { k_ffma_const<<<nb, 256>>>(d_out, d_cyc, t, 0.5f); }, d_cyc);
    run("FFMA  tap=reg (3 reg operands)", n1, bps, nsm, [&](int nb) { k_ffma_reg<<<nb, 256>>>(d_out, d_cyc, d_in, 0.5f); }, d_cyc);
    run("FFMA2 tap=const scalar", 2 * n1, bps, nsm, [&](int nb) { k_ffma2_const<<<nb, 256>>>(d_out, d_cyc, t, 0.5f); }, d_cyc);
    run("FFMA2 tap=reg pair", 2 * n1, bps, nsm, [&](int nb) { k_ffma2_reg<<<nb, 256>>>(d_out, d_cyc, d_in, 0.5f); }, d_cyc);
    run("FFMA x2 + LDS.128 every 8", 2 * n1, bps, nsm, [&](int nb) { k_mix<4, false><<<nb, 256>>>(d_out, d_cyc, t, 0.5f); }, d_cyc);
    run("FFMA2   + LDS.128 every 8", 2 * n1, bps, nsm, [&](int nb) { k_mix<4, true><<<nb, 256>>>(d_out, d_cyc, t, 0.5f); }, d_cyc);
    run("FFMA x2 + LDS.128 every 4", 2 * n1, bps, nsm, [&](int nb) { k_mix<2, false><<<nb, 256>>>(d_out, d_cyc, t, 0.5f); }, d_cyc);
    run("FFMA2   + LDS.128 every 4", 2 * n1, bps, nsm, [&](int nb) { k_mix<2, true><<<nb, 256>>>(d_out, d_cyc, t, 0.5f); }, d_cyc);
  }
  for (int bps : {2, 4}) {
    run("MUFU rsqrtf", n1, bps, nsm, [&](int nb) { k_mufu<0><<<nb, 256>>>(d_out, d_cyc, 0.5f); }, d_cyc);
    run("MUFU __log2f", n1, bps, nsm, [&](int nb) { k_mufu<1><<<nb, 256>>>(d_out, d_cyc, 0.5f); }, d_cyc);
    run("logf (precise)", n1, bps, nsm, [&](int nb) { k_mufu<2><<<nb, 256>>>(d_out, d_cyc, 0.5f); }, d_cyc);
    run("sqrtf (IEEE)", n1, bps, nsm, [&](int nb) { k_mufu<3><<<nb, 256>>>(d_out, d_cyc, 0.5f); }, d_cyc);
    run("1/x (IEEE)", n1, bps, nsm, [&](int nb) { k_mufu<4><<<nb, 256>>>(d_out, d_cyc, 0.5f); }, d_cyc);
    run("__fdividef", n1, bps, nsm, [&](int nb) { k_mufu<5><<<nb, 256>>>(d_out, d_cyc, 0.5f); }, d_cyc);
  }
  printf("done\n");
  return 0;
}
