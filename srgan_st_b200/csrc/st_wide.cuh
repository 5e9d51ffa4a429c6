// "Wide-strip" structure-tensor kernels for sm_100a: one big tile per CTA, one CTA per SM.
//
// Same mathematics and C ABI as st_kernels.cuh (loss.py:399-413, utils.py:212-279); a different
// decomposition aimed at the 96-pixel training crops of train.py / warmup.py, where the tiled
// kernels lose most of their time to halo recomputation, load latency and the V-plane round trip:
//
//   * tile = TH x TW = 48 x 96 pixels (half a training crop: no x halo inside the image at all),
//     NT = 384 threads, ~166 KB of shared memory;
//   * smoothing runs HORIZONTAL FIRST, in scatter form (every Ix/Iy row-pair value is loaded once,
//     the three products are formed in registers and fed to 8 output columns x 3 channels), then
//     VERTICAL with one lane per column and RS rows, which leaves the smoothed tensor directly in
//     registers for the per-pixel chain -- the vertical-pass plane of the tiled kernel is never
//     written or re-read;
//   * the HR gray tile is prefetched in three register slices while the SR tile is being
//     filtered, so only the first load of a CTA is exposed.
//
// Shared-memory planes are row-pair interleaved like in st_kernels.cuh and all filter passes are
// packed FFMA2 with taps from the constant bank.
#pragma once
#include "st_kernels.cuh"

namespace srst {

// Debug time stamps (SRST_ST_DEBUG=1): lane 0 of every warp of CTA 0 records clock64() at slot k.
#ifdef SRST_EMULATE
#define SRST_STAMP(dbg, k) ((void)0)
#define SRST_GSTAMP(dbg, k) ((void)0)
#else
#define SRST_STAMP(dbg, k) do { if ((dbg) && blockIdx.x == 0 && (threadIdx.x & 31) == 0) (dbg)[(threadIdx.x >> 5) * 32 + (k)] = clock64(); } while (0)
// wall-clock (ns) stamps of warp 0 of CTA 0 and of the last CTA, slots 24..31
#define SRST_GSTAMP(dbg, k) do { if ((dbg) && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); (dbg)[(blockIdx.x == 0 ? 0 : 32) + (k)] = (long long)t_; } } while (0)
#endif

template <int TH_, int TW_, int RS_, int NT_>
struct StWideFwdCfg {
  static constexpr int RG = 2, RK = 8;
  static constexpr int TH = TH_, TW = TW_, RS = RS_, NT = NT_;
  static constexpr int OFF = 4, HXD = 8, HXG = HXD + OFF;                              // x halos (multiples of 4)
  static constexpr int GH = TH + 2 * (RG + RK), GW = TW + 2 * HXG, PG = smem_pitch(2 * GW);  // gray region
  static constexpr int DH = TH + 2 * RK, DW = TW + 2 * HXD, PD = smem_pitch(2 * DW);         // Ix, Iy region
  static constexpr int PH = smem_pitch(2 * TW);                                        // H planes: DH rows x TW cols
  static constexpr int G_FLOATS = (GH / 2) * PG, D_FLOATS = (DH / 2) * PD, H_FLOATS = (DH / 2) * PH;
  // smem: D0 | D1 | H[3] (the SR gray tile aliases the start of H: it is dead before H is written) | HR gray tile
  static constexpr int H_OFF = 2 * D_FLOATS, G2_OFF = H_OFF + 3 * H_FLOATS;
  static constexpr int SMEM_FLOATS = G2_OFF + G_FLOATS;
  static constexpr size_t SMEM_BYTES = sizeof(float) * SMEM_FLOATS;
  static constexpr int NSEG = TH / RS;
  static constexpr int NGI = (GH / 2) * (GW / 4);       // gray-tile items (row pair x 4 columns)
  static constexpr int NSL = (NGI + NT - 1) / NT;       // register slices of one gray tile
  static constexpr int BW_LO = (OFF - RG) / 2 * 2, BW_HI = (OFF + 4 + RG + 1) / 2 * 2, BWIN = BW_HI - BW_LO;
  static_assert(G_FLOATS <= 3 * H_FLOATS, "SR gray tile must fit inside the H planes");
  static_assert(TH % RS == 0 && RS % 2 == 0 && TW % 8 == 0 && NT % 32 == 0 && NT <= 1024, "bad wide tile");
  static_assert(SMEM_BYTES <= 227 * 1024, "wide tile does not fit in shared memory");
};

// One gray-tile item (row pair x 4 columns) held in registers between issue and commit.
struct GrayItemRegs {
  float4 v[2][3];
  bool ok[2];
};

template <class C>
SRST_DEV void gray_item_issue(GrayItemRegs& r, const float* __restrict__ base, int H, int W, int gy0, int gx0, int it) {
  constexpr int C4 = C::GW / 4;
  const size_t plane = (size_t)H * W;
  const int q = it / C4, c4 = it - q * C4;
  const int gx = gx0 + 4 * c4;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int gy = gy0 + 2 * q + hf;
    r.ok[hf] = (it < C::NGI) && gy >= 0 && gy < H && gx >= 0 && gx < W;  // W % 4 == 0: all-in or all-out
    if (r.ok[hf]) {
      const float* p = base + (size_t)gy * W + gx;
      r.v[hf][0] = ldg4(p); r.v[hf][1] = ldg4(p + plane); r.v[hf][2] = ldg4(p + 2 * plane);
    }
  }
}

template <class C>
SRST_DEV void gray_item_commit(const GrayItemRegs& r, float* sG, int it) {
  constexpr int C4 = C::GW / 4;
  if (it >= C::NGI) return;
  const int q = it / C4, c4 = it - q * C4;
  float v[2][4];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    if (r.ok[hf]) {
      v[hf][0] = gray_of(r.v[hf][0].x, r.v[hf][1].x, r.v[hf][2].x);
      v[hf][1] = gray_of(r.v[hf][0].y, r.v[hf][1].y, r.v[hf][2].y);
      v[hf][2] = gray_of(r.v[hf][0].z, r.v[hf][1].z, r.v[hf][2].z);
      v[hf][3] = gray_of(r.v[hf][0].w, r.v[hf][1].w, r.v[hf][2].w);
    } else {
      v[hf][0] = v[hf][1] = v[hf][2] = v[hf][3] = 0.f;
    }
  }
  float* o = sG + q * C::PG + 8 * c4;
  st4(o, make_float4(v[0][0], v[1][0], v[0][1], v[1][1]));
  st4(o + 4, make_float4(v[0][2], v[1][2], v[0][3], v[1][3]));
}

// Horizontal rho-pass of the three products, scatter form: outputs cols [8cg, 8cg+8) of row pair q
// of the H planes from D columns [8cg, 8cg+24).  408 FFMA2 + 72 FMUL2 per 24 LDS.128.
template <class C, class Taps>
SRST_DEV void wide_hpass_item(const float* sD0, const float* sD1, float* sH, int q, int cg, const Taps& tp) {
  float2 acc[3][8];
#pragma unroll
  for (int o = 0; o < 8; ++o) { acc[0][o] = make_float2(0.f, 0.f); acc[1][o] = make_float2(0.f, 0.f); acc[2][o] = make_float2(0.f, 0.f); }
  const float* p0 = sD0 + q * C::PD + 16 * cg;
  const float* p1 = sD1 + q * C::PD + 16 * cg;
#pragma unroll
  for (int m = 0; m < 4 + C::RK; ++m) {  // two columns per LDS.128
    const float4 a = ld4(p0 + 4 * m), b = ld4(p1 + 4 * m);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * m + h;
      const float2 ix = h ? make_float2(a.z, a.w) : make_float2(a.x, a.y);
      const float2 iy = h ? make_float2(b.z, b.w) : make_float2(b.x, b.y);
      const float2 pxx = mul2(ix, ix), pyy = mul2(iy, iy), pxy = mul2(ix, iy);
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int t = j - o;
        if (t >= 0 && t <= 2 * C::RK) {
          acc[0][o] = ffma2(pxx, bcast2(tp.k[t]), acc[0][o]);
          acc[1][o] = ffma2(pyy, bcast2(tp.k[t]), acc[1][o]);
          acc[2][o] = ffma2(pxy, bcast2(tp.k[t]), acc[2][o]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float* o = sH + c * C::H_FLOATS + q * C::PH + 16 * cg;
#pragma unroll
    for (int m = 0; m < 4; ++m) st4(o + 4 * m, make_float4(acc[c][2 * m].x, acc[c][2 * m].y, acc[c][2 * m + 1].x, acc[c][2 * m + 1].y));
  }
}

// Phases B-D of one image: gray tile in sG -> smoothed tensor of this thread's column / row segment
// in S[3][RS/2] (.x = even row, .y = odd row).  PF: prefetch the other image's gray tile into sG2
// in register slices hidden behind the three phases.
template <class C, bool PF, class Taps>
SRST_DEV void wide_unit(float* smem, const float* sG, const float* __restrict__ pf_base, float* sG2, int H, int W, int y0,
                        int x0, const Taps& tp, int tid, float2 (&S)[3][C::RS / 2], long long* dbg, int kb) {
  float* sD0 = smem;
  float* sD1 = smem + C::D_FLOATS;
  float* sH = smem + C::H_OFF;
  const int gy0 = y0 - (C::RG + C::RK), gx0 = x0 - C::HXG;
  GrayItemRegs pf;

  // Phase B: Ix, Iy on the D region, zero outside the image (the reference zero-pads the products)
  if (PF) gray_item_issue<C>(pf, pf_base, H, W, gy0, gx0, tid);
  for (int it = tid; it < (C::DH / 2) * (C::DW / 4); it += C::NT) {
    const int seg = it / (C::DH / 2), q = it - seg * (C::DH / 2);
    const int dx0 = 4 * seg;
    const int gy = y0 - C::RK + 2 * q, gx = x0 - C::HXD + dx0;
    float2 Ix[4], Iy[4];
    if (gy + 1 >= 0 && gy < H && gx + 3 >= 0 && gx < W) {
      const float* p = sG + q * C::PG + 2 * (dx0 + C::BW_LO);
      grad_rowpair<C::RG, 4, C::BWIN, C::OFF - C::BW_LO, C::PG, true>(p, p, tp, Ix, Iy);
      const bool r0 = gy >= 0, r1 = gy + 1 < H;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (gx + j >= 0) && (gx + j < W);
        Ix[j].x = (ok && r0) ? Ix[j].x : 0.f;
        Ix[j].y = (ok && r1) ? Ix[j].y : 0.f;
        Iy[j].x = (ok && r0) ? Iy[j].x : 0.f;
        Iy[j].y = (ok && r1) ? Iy[j].y : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { Ix[j] = make_float2(0.f, 0.f); Iy[j] = make_float2(0.f, 0.f); }
    }
    float* o0 = sD0 + q * C::PD + 2 * dx0;
    float* o1 = sD1 + q * C::PD + 2 * dx0;
    st4(o0, make_float4(Ix[0].x, Ix[0].y, Ix[1].x, Ix[1].y));
    st4(o0 + 4, make_float4(Ix[2].x, Ix[2].y, Ix[3].x, Ix[3].y));
    st4(o1, make_float4(Iy[0].x, Iy[0].y, Iy[1].x, Iy[1].y));
    st4(o1 + 4, make_float4(Iy[2].x, Iy[2].y, Iy[3].x, Iy[3].y));
  }
  SRST_STAMP(dbg, kb + 0);
  if (PF) gray_item_commit<C>(pf, sG2, tid);
  SRST_STAMP(dbg, kb + 1);
  __syncthreads();
  SRST_STAMP(dbg, kb + 2);

  // Phase C: horizontal rho-pass (H planes alias the SR gray tile, which is dead from here on)
  if (PF && C::NSL > 1) gray_item_issue<C>(pf, pf_base, H, W, gy0, gx0, C::NT + tid);
  for (int it = tid; it < (C::DH / 2) * (C::TW / 8); it += C::NT) {
    const int cg = it / (C::DH / 2), q = it - cg * (C::DH / 2);
    const int gy = y0 - C::RK + 2 * q;
    if (gy + 1 >= 0 && gy < H && x0 + 8 * cg < W) {
      wide_hpass_item<C>(sD0, sD1, sH, q, cg, tp);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float* o = sH + c * C::H_FLOATS + q * C::PH + 16 * cg;
#pragma unroll
        for (int m = 0; m < 4; ++m) st4(o + 4 * m, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
  }
  SRST_STAMP(dbg, kb + 3);
  if (PF && C::NSL > 1) gray_item_commit<C>(pf, sG2, C::NT + tid);
  SRST_STAMP(dbg, kb + 4);
  __syncthreads();
  SRST_STAMP(dbg, kb + 5);

  // Phase D: vertical rho-pass; a lane owns one column and RS output rows, results stay in registers
  if (PF && C::NSL > 2) gray_item_issue<C>(pf, pf_base, H, W, gy0, gx0, 2 * C::NT + tid);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < C::RS / 2; ++j) S[c][j] = make_float2(0.f, 0.f);
  if (tid < C::TW * C::NSEG) {
    const int seg = tid / C::TW, hc = tid - seg * C::TW;
    if (x0 + hc < W && y0 + seg * C::RS < H) {
      const float* p = sH + (seg * (C::RS / 2)) * C::PH + 2 * hc;
#pragma unroll
      for (int rq = 0; rq < C::RS / 2 + C::RK; ++rq) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float2 v = ld2(p + c * C::H_FLOATS + rq * C::PH);
#pragma unroll
          for (int jp = 0; jp < C::RS / 2; ++jp) {
            const int u0 = 2 * rq - 2 * jp;  // tap-pair index of input row 2rq for output pair jp
            if (u0 >= 0 && u0 <= 2 * C::RK + 1) S[c][jp] = ffma2(bcast2(v.x), tp.kp[u0], S[c][jp]);
            if (u0 + 1 >= 0 && u0 + 1 <= 2 * C::RK + 1) S[c][jp] = ffma2(bcast2(v.y), tp.kp[u0 + 1], S[c][jp]);
          }
        }
      }
    }
  }
  SRST_STAMP(dbg, kb + 6);
  if (PF && C::NSL > 2) gray_item_commit<C>(pf, sG2, 2 * C::NT + tid);
  SRST_STAMP(dbg, kb + 7);
  if (PF) {
    for (int sl = 3; sl < C::NSL; ++sl) {  // tiles with more than three slices: the rest is loaded in place
      gray_item_issue<C>(pf, pf_base, H, W, gy0, gx0, sl * C::NT + tid);
      gray_item_commit<C>(pf, sG2, sl * C::NT + tid);
    }
    __syncthreads();  // H reads done before the next image's phase C; HR gray tile complete
  }
  SRST_STAMP(dbg, kb + 8);
}

// Per-pixel chain of one thread's NP row pairs, branch-free so that the scheduler can interleave
// the long SFU/FMA dependency chains of different pairs.  Returns the masked sum of distances.
template <bool WANT_HR, int NP>
SRST_DEV float wide_chain(const float2 (&S1)[3][NP], const float2 (&S2)[3][NP], bool norm, float eps, int H, int W, int gy0,
                          int gx, StPixelGrad2 (&G)[NP]) {
  float lsum = 0.f;
#pragma unroll
  for (int jp = 0; jp < NP; ++jp) {
    G[jp].da = G[jp].db = G[jp].dc = G[jp].de = G[jp].df = G[jp].dh = make_float2(0.f, 0.f);
    const float2 d = st_pixel2<true, WANT_HR>(S1[0][jp], S1[1][jp], S1[2][jp], S2[0][jp], S2[1][jp], S2[2][jp], norm, eps, G[jp]);
    const int gy = gy0 + 2 * jp;
    lsum += (gx < W && gy < H) ? d.x : 0.f;
    lsum += (gx < W && gy + 1 < H) ? d.y : 0.f;
  }
  return lsum;
}

// Masked stores of the ds planes of one thread's column (gx) and NP row pairs from row gy0.
template <bool WANT_HR, int NP>
SRST_DEV void wide_store(const StPixelGrad2 (&G)[NP], float* __restrict__ ds_sr, float* __restrict__ ds_hr, size_t img_off,
                         int H, int W, int gy0, int gx) {
  if (gx >= W) return;
  const size_t plane = (size_t)H * W;
#pragma unroll
  for (int jp = 0; jp < NP; ++jp) {
    const int gy = gy0 + 2 * jp;
    const bool r0 = gy < H, r1 = gy + 1 < H;
    const size_t o = img_off + (size_t)gy * W + gx;
    if (ds_sr) {
      if (r0) { ds_sr[o] = G[jp].da.x; ds_sr[o + plane] = G[jp].db.x; ds_sr[o + 2 * plane] = G[jp].dc.x; }
      if (r1) { ds_sr[o + W] = G[jp].da.y; ds_sr[o + plane + W] = G[jp].db.y; ds_sr[o + 2 * plane + W] = G[jp].dc.y; }
    }
    if (WANT_HR) {
      if (r0) { ds_hr[o] = G[jp].de.x; ds_hr[o + plane] = G[jp].df.x; ds_hr[o + 2 * plane] = G[jp].dh.x; }
      if (r1) { ds_hr[o + W] = G[jp].de.y; ds_hr[o + plane + W] = G[jp].df.y; ds_hr[o + 2 * plane + W] = G[jp].dh.y; }
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
st_wide_forward_kernel(const __grid_constant__ StFwdParams<C::RG, C::RK> P) {
  SRST_DYN_SMEM(float, smem);
  __shared__ float s_red[32];
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int tx = t % P.tiles_x;
  t /= P.tiles_x;
  const int ty = t % P.tiles_y;
  const int b = t / P.tiles_y;
  const int y0 = ty * C::TH, x0 = tx * C::TW;
  const int H = P.H, W = P.W;
  const size_t img_off = (size_t)b * 3 * H * W;
  float* sG1 = smem + C::H_OFF;   // SR gray tile (aliases the H planes)
  float* sG2 = smem + C::G2_OFF;  // HR gray tile

  SRST_GSTAMP(P.debug, 24);
  pdl_wait();     // the previous kernel of the stream (producer of sr / user of the workspace) is done
  SRST_GSTAMP(P.debug, 25);
  pdl_trigger();  // the next PDL-launched kernel may begin its launch; it waits for this grid to finish
  SRST_STAMP(P.debug, 0);
  // Phase A: SR gray tile, every item of the thread in flight at once
  load_gray_tile<C::GH, C::GW, C::PG, C::NT, C::NSL>(sG1, P.sr + img_off, H, W, y0 - (C::RG + C::RK), x0 - C::HXG, true, tid);
  SRST_STAMP(P.debug, 1);
  __syncthreads();
  SRST_STAMP(P.debug, 2);

  float2 S1[3][C::RS / 2], S2[3][C::RS / 2];
  wide_unit<C, true>(smem, sG1, P.hr + img_off, sG2, H, W, y0, x0, P.taps, tid, S1, P.debug, 3);
  wide_unit<C, false>(smem, sG2, nullptr, nullptr, H, W, y0, x0, P.taps, tid, S2, P.debug, 12);

  // Per-pixel chain; the block's loss partial is published BEFORE the ds stores are issued, so the
  // fence in front of the ticket does not wait for them and the ticket round trip overlaps them.
  const bool want_hr = P.ds_hr != nullptr;
  const int seg = tid / C::TW, hc = tid - seg * C::TW;
  const bool active = tid < C::TW * C::NSEG;
  const int gy0 = y0 + seg * C::RS, gx = active ? x0 + hc : W;
  StPixelGrad2 G[C::RS / 2];
  float lsum;
  if (want_hr) lsum = wide_chain<true, C::RS / 2>(S1, S2, P.normalize != 0, P.eps, H, W, gy0, gx, G);
  else lsum = wide_chain<false, C::RS / 2>(S1, S2, P.normalize != 0, P.eps, H, W, gy0, gx, G);
  SRST_STAMP(P.debug, 21);

  // Deterministic loss reduction: block partial -> workspace; the last block to take a ticket sums
  // all partials in a fixed order (double) and re-zeroes the workspace for the next call.
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0) s_red[tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) {
    float bs = 0.f;
    for (int w = 0; w < C::NT / 32; ++w) bs += s_red[w];
    P.partials[blockIdx.x] = bs;
    __threadfence();
    const unsigned int tk = atomicAdd(P.ticket, 1u);
    s_last = (tk == gridDim.x - 1) ? 1u : 0u;
  }
  if (want_hr) wide_store<true, C::RS / 2>(G, P.ds_sr, P.ds_hr, img_off, H, W, gy0, gx);
  else wide_store<false, C::RS / 2>(G, P.ds_sr, P.ds_hr, img_off, H, W, gy0, gx);
  SRST_STAMP(P.debug, 22);
  SRST_GSTAMP(P.debug, 26);
  if (tid >= 32) return;
  __syncwarp();
  if (s_last) {
    __threadfence();
    double acc = 0.0;
    for (unsigned int i = tid; i < gridDim.x; i += 32) {
      acc += (double)__ldcg(P.partials + i);
      P.partials[i] = 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (tid == 0) {
      P.loss_out[0] = (float)(acc * (double)P.inv_count);
      *P.ticket = 0u;
    }
  }
  SRST_STAMP(P.debug, 23);
  SRST_GSTAMP(P.debug, 27);
}

}  // namespace srst
