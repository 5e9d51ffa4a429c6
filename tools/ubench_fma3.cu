// Micro-benchmark #3: register-only FFMA2/FMUL2/FADD2 forms, shared-memory stream rates, barrier cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fma3 tools/ubench_fma3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 1024, NACC = 12;

// FORM 0: acc2 = a2 * b2 + acc2 (all register pairs, distinct)      FFMA2 R,R,R
// FORM 1: acc2 = acc2 * b2  (FMUL2 R,R)     FORM 2: acc2 = acc2 + b2 (FADD2 R,R)
// FORM 3: scalar x2: acc.x = a.x*b.x+acc.x; acc.y = a.y*b.y+acc.y   (2 FFMA R,R,R)
// FORM 4: acc2 = a2 * bcast(reg scalar) + acc2                       FFMA2 R, R.F32?, R
template <int FORM>
__global__ void __launch_bounds__(256) kern(float* out, const float* in, float x0) {
  float2 acc[NACC], a[4], b[4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = make_float2(in[2 * i] + x0 + threadIdx.x * 1e-6f, in[2 * i + 1] - x0 - threadIdx.x * 1e-6f);
    b[i] = make_float2(in[8 + 2 * i] * x0 + threadIdx.x * 2e-6f, in[9 + 2 * i] + x0 - threadIdx.x * 3e-6f);
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (FORM == 0) acc[i] = __ffma2_rn(a[i & 3], b[(i >> 2) & 3], acc[i]);
      if (FORM == 1) acc[i] = __fmul2_rn(acc[i], b[i & 3]);
      if (FORM == 2) acc[i] = __fadd2_rn(acc[i], b[i & 3]);
      if (FORM == 3) { acc[i].x = fmaf(a[i & 3].x, b[(i >> 2) & 3].x, acc[i].x); acc[i].y = fmaf(a[i & 3].y, b[(i >> 2) & 3].y, acc[i].y); }
      if (FORM == 4) acc[i] = __ffma2_rn(a[i & 3], make_float2(b[(i >> 2) & 3].x, b[(i >> 2) & 3].x), acc[i]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Shared-memory stream: each warp reads consecutive conflict-free vectors.  MODE 0: LDS.32, 1: LDS.64, 2: LDS.128,
// 3: STS.64, 4: LDS.64 + 8 FFMA2(const-free reg form A) per load, 5: LDS.128 + 8 FFMA2 per load
template <int MODE>
__global__ void __launch_bounds__(256) smem_kern(float* out, float x0) {
  extern __shared__ float4 sm4[];
  float* sm = reinterpret_cast<float*>(sm4);
  for (int i = threadIdx.x; i < 8192; i += 256) sm[i] = x0 * i;
  __syncthreads();
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(i, threadIdx.x);
  float4 s4 = make_float4(0, 0, 0, 0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int base = ((it * 8 + u) * 37 + w * 64) & 1023;  // float4 index window of 1024 float4 = 16 KB .. within 32 KB
      if (MODE == 0) { float v = sm[(base * 4 + lane) & 8191]; s4.x += v; }
      if (MODE == 1) { float2 v = *reinterpret_cast<float2*>(&sm[((base * 4) + 2 * lane) & 8191]); s4.x += v.x; s4.y += v.y; }
      if (MODE == 2) { float4 v = sm4[(base + lane) & 2047]; s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w; }
      if (MODE == 3) { *reinterpret_cast<float2*>(&sm[((base * 4) + 2 * lane) & 8191]) = make_float2(s4.x, (float)u); }
      if (MODE == 4) {
        float2 v = *reinterpret_cast<float2*>(&sm[((base * 4) + 2 * lane) & 8191]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(v, make_float2(x0, x0), acc[i]);
      }
      if (MODE == 5) {
        float4 v = sm4[(base + lane) & 2047];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i] = __ffma2_rn(make_float2(v.x, v.y), make_float2(x0, x0), acc[i]); acc[i + 4] = __ffma2_rn(make_float2(v.z, v.w), make_float2(x0, x0), acc[i + 4]); }
      }
    }
  }
  float s = s4.x + s4.y + s4.z + s4.w;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Barrier cost: NT threads loop over bar.sync with a tiny amount of work in between.
__global__ void bar_kern(float* out, long long* cyc) {
  float a = threadIdx.x;
  long long c0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    a = fmaf(a, 1.0001f, 0.5f);
    __syncthreads();
  }
  long long c1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

int main() {
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  float* d_out; float* d_in; long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(float) * 1024 * 8 * nsm)); CK(cudaMalloc(&d_in, 64 * sizeof(float))); CK(cudaMalloc(&d_cyc, 8 * nsm * sizeof(long long)));
  std::vector<float> h(64); for (int i = 0; i < 64; ++i) h[i] = 1e-3f * (i + 1);
  CK(cudaMemcpy(d_in, h.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[5] = {"FFMA2 R,R,R (pairs)        ", "FMUL2 R,R                  ", "FADD2 R,R                  ", "2x FFMA R,R,R (scalar)     ", "FFMA2 R, bcast(R), R       "};
  for (int bps : {1, 2, 4}) for (int f = 0; f < 5; ++f) {
    int nb = bps * nsm;
    auto launch = [&]() { if (f == 0) kern<0><<<nb, 256>>>(d_out, d_in, 0.5f); if (f == 1) kern<1><<<nb, 256>>>(d_out, d_in, 0.5f); if (f == 2) kern<2><<<nb, 256>>>(d_out, d_in, 0.5f);
                          if (f == 3) kern<3><<<nb, 256>>>(d_out, d_in, 0.5f); if (f == 4) kern<4><<<nb, 256>>>(d_out, d_in, 0.5f); };
    launch(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane = 2.0 * ITERS * NACC * 256.0 * nb;
    printf("%s blocks/SM=%d  %.3f ms  %.1f lane-op/clk/SM\n", names[f], bps, ms, lane / (ms * 1e-3) / 1.965e9 / nsm);
  }
  const char* sn[6] = {"LDS.32 stream ", "LDS.64 stream ", "LDS.128 stream", "STS.64 stream ", "LDS.64 + 8 FFMA2 ", "LDS.128 + 8 FFMA2"};
  const int bytes_per[6] = {4, 8, 16, 8, 8, 16};
  for (int bps : {1, 2, 4}) for (int f = 0; f < 6; ++f) {
    int nb = bps * nsm;
    auto launch = [&]() { if (f == 0) smem_kern<0><<<nb, 256, 32768>>>(d_out, 0.5f); if (f == 1) smem_kern<1><<<nb, 256, 32768>>>(d_out, 0.5f); if (f == 2) smem_kern<2><<<nb, 256, 32768>>>(d_out, 0.5f);
                          if (f == 3) smem_kern<3><<<nb, 256, 32768>>>(d_out, 0.5f); if (f == 4) smem_kern<4><<<nb, 256, 32768>>>(d_out, 0.5f); if (f == 5) smem_kern<5><<<nb, 256, 32768>>>(d_out, 0.5f); };
    launch(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)ITERS * 8 * 256.0 * nb * bytes_per[f];
    double lane = (f >= 4) ? 2.0 * ITERS * 8 * 8 * 256.0 * nb : 0.0;
    printf("%s blocks/SM=%d  %.3f ms  %.1f B/clk/SM  %.1f lane-FMA/clk/SM\n", sn[f], bps, ms, bytes / (ms * 1e-3) / 1.965e9 / nsm, lane / (ms * 1e-3) / 1.965e9 / nsm);
  }
  for (int nt : {128, 256, 512, 1024}) {
    bar_kern<<<nsm, nt>>>(d_out, d_cyc); CK(cudaDeviceSynchronize());
    std::vector<long long> c(nsm); CK(cudaMemcpy(c.data(), d_cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
    printf("bar.sync %4d threads: %.1f cycles per iteration\n", nt, (double)c[0] / ITERS);
  }
  return 0;
}
