// Streaming (row-marching) forward kernel of the structure-tensor loss for sm_100a.
//
// The tiled kernel in st_kernels.cuh runs its phases one after another and pays a barrier, a
// partially filled round and a recomputed vertical halo at every phase.  This kernel turns the
// phases into a SOFTWARE PIPELINE over 16-row slots of a column strip, with one warp role per
// stage and a single CTA barrier per slot:
//
//      interval t :   LG loads+grays slot t | GR differentiates slot t-1 | VS smooths slot t-2
//                     vertically            | HC smooths slot t-3 horizontally + per-pixel chain
//
// Stages hand rows over through ring buffers in shared memory (gray: 3 slots, Ix/Iy: 3 slots,
// V: 2 slots), so within an interval no role reads what another writes.  A persistent CTA walks a
// list of units (image, 64-column strip, row segment) back to back, so the pipeline never drains
// between units, every global load is issued a full interval before it is needed, and rows are
// differentiated / vertically smoothed exactly once per strip (no vertical halo recompute).
// All filter arithmetic is the same row-pair FFMA2 code as the tiled kernel.
#pragma once
#include "st_kernels.cuh"

namespace srst {

template <int TW_, int WLG_, int WGR_, int WVS_, int WHC_>
struct StStreamCfg {
  static constexpr int TW = TW_, RG = 2, RK = 8;
  static constexpr int RB = 16, RP = RB / 2;  // rows / row pairs per slot
  static constexpr int WLG = WLG_, WGR = WGR_, WVS = WVS_, WHC = WHC_;
  static constexpr int NT = 32 * (WLG + WGR + WVS + WHC);
  static constexpr int HXD = 8, OFF = 4, HXG = HXD + OFF;
  static constexpr int GW = TW + 2 * HXG, DW = TW + 2 * HXD;
  static constexpr int PG = smem_pitch(2 * GW), PD = smem_pitch(2 * DW), PV = PD;
  static constexpr int G_FLOATS = 3 * RP * PG;  // per image
  static constexpr int D_FLOATS = 3 * RP * PD;  // per image and plane
  static constexpr int V_FLOATS = 2 * RP * PV;  // per image and channel
  static constexpr int STAGE_FLOATS = 2 * 3 * RB * GW;
  static constexpr int SMEM_FLOATS = 2 * G_FLOATS + 4 * D_FLOATS + 6 * V_FLOATS + STAGE_FLOATS;
  static constexpr size_t SMEM_BYTES = sizeof(float) * SMEM_FLOATS;
  static constexpr int BW_LO = (OFF - RG) / 2 * 2, BWIN = (OFF + 4 + RG + 1) / 2 * 2 - BW_LO;  // gradient window
  static constexpr int HWIN = (HXD + 2 + RK + 1) / 2 * 2;                                       // horizontal window (2 outputs)
  static constexpr int LG_IPT = (2 * RP * (GW / 4) + 32 * WLG - 1) / (32 * WLG);                // loader items per thread
  static_assert(TW % 4 == 0 && NT <= 1024, "bad stream configuration");
};

constexpr int kBarStreamLG = 6;

struct StStreamParams {
  const float* sr;
  const float* hr;
  float* ds_sr;
  float* ds_hr;
  float* partials;
  unsigned int* ticket;
  float* loss_out;
  int B, H, W;
  int nstrips, nsegs, segh, nunits;
  int normalize;
  float eps;
  float inv_count;
  long long* debug;  // optional: per-warp (work, wait) cycle counters of CTA 0
  StTaps<2, 8> taps;
};

// Position of one role in the CTA's slot stream.
struct StStreamCursor {
  int unit, k, nslots, S;
  int b, x0, ya, yb;
  SRST_DEV void decode(const StStreamParams& P, int tw) {
    const int sg = unit % P.nsegs;
    const int r = unit / P.nsegs;
    const int sx = r % P.nstrips;
    b = r / P.nstrips;
    x0 = sx * tw;
    ya = sg * P.segh;
    yb = min(P.H, ya + P.segh);
    nslots = (yb - ya + 15) / 16 + 2;
  }
  SRST_DEV void init(const StStreamParams& P, int tw, int first_unit) {
    unit = first_unit; k = 0; S = 0;
    if (unit < P.nunits) decode(P, tw);
  }
  SRST_DEV void advance(const StStreamParams& P, int tw, int stride) {
    ++k; ++S;
    if (k == nslots) {
      unit += stride; k = 0;
      if (unit < P.nunits) decode(P, tw);
    }
  }
};

// LG helpers.  The raw RGB rows of a slot ([img][plane][16 rows][GW]) are staged with 16-byte
// cp.async copies (zero-filled outside the image) one interval before they are converted to gray,
// so no register is held while the data is in flight and the loader never waits on HBM.
template <class C>
SRST_DEV void lg_issue(float* sStage, const StStreamParams& P, const StStreamCursor& c, int rtid, int rthreads) {
  constexpr int C4 = C::GW / 4;
  constexpr int ROWS = 2 * 3 * C::RB;  // (img, plane, row) flattened
  const size_t plane = (size_t)P.H * P.W;
  const size_t img_off = (size_t)c.b * 3 * plane;
  const int gyb = c.ya - 20 + 16 * c.k, gxb = c.x0 - C::HXG;
  for (int it = rtid; it < ROWS * C4; it += rthreads) {
    const int c4 = it % C4, rr = it / C4;
    const int r = rr % C::RB, ip = rr / C::RB;  // ip = img * 3 + plane
    const int img = ip / 3, pl = ip - 3 * img;
    const int gy = gyb + r, gx = gxb + 4 * c4;
    const bool ok = gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
    const float* base = (img ? P.hr : P.sr) + img_off;
    cp_async16(sStage + rr * C::GW + 4 * c4, ok ? base + pl * plane + (size_t)gy * P.W + gx : base, ok);
  }
  cp_async_commit();
}

template <class C>
SRST_DEV void lg_commit(float* sG, const float* sStage, const StStreamCursor& c, int rtid, int rthreads) {
  constexpr int C4 = C::GW / 4;
  const int pbase = (c.S % 3) * C::RP;
  for (int it = rtid; it < 2 * C::RP * C4; it += rthreads) {
    const int img = it / (C::RP * C4), rem = it - img * (C::RP * C4);
    const int q = rem / C4, c4 = rem - q * C4;
    float v[2][4];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float* p = sStage + ((img * 3) * C::RB + 2 * q + hf) * C::GW + 4 * c4;
      const float4 R = ld4(p), G = ld4(p + C::RB * C::GW), B = ld4(p + 2 * C::RB * C::GW);
      v[hf][0] = gray_of(R.x, G.x, B.x);
      v[hf][1] = gray_of(R.y, G.y, B.y);
      v[hf][2] = gray_of(R.z, G.z, B.z);
      v[hf][3] = gray_of(R.w, G.w, B.w);
    }
    float* o = sG + img * C::G_FLOATS + (pbase + q) * C::PG + 8 * c4;
    st4(o, make_float4(v[0][0], v[1][0], v[0][1], v[1][1]));
    st4(o + 4, make_float4(v[0][2], v[1][2], v[0][3], v[1][3]));
  }
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
st_stream_forward_kernel(const __grid_constant__ StStreamParams P) {
  SRST_DYN_SMEM(float, smem);
  __shared__ float s_red[32];
  __shared__ unsigned int s_last;
  float* sG = smem;                         // [img][3*RP][PG]
  float* sD = sG + 2 * C::G_FLOATS;         // [img][plane][3*RP][PD]
  float* sV = sD + 4 * C::D_FLOATS;         // [img][ch][2*RP][PV]
  float* sStage = sV + 6 * C::V_FLOATS;     // [img][plane][RB][GW] raw RGB of the slot being prefetched
  const auto& tp = P.taps;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int H = P.H, W = P.W;
  const size_t plane = (size_t)H * W;

  // role of this warp and its lag behind the loader
  int role, rtid, rthreads;
  if (warp < C::WHC) { role = 3; rtid = tid; rthreads = 32 * C::WHC; }
  else if (warp < C::WHC + C::WVS) { role = 2; rtid = tid - 32 * C::WHC; rthreads = 32 * C::WVS; }
  else if (warp < C::WHC + C::WVS + C::WGR) { role = 1; rtid = tid - 32 * (C::WHC + C::WVS); rthreads = 32 * C::WGR; }
  else { role = 0; rtid = tid - 32 * (C::WHC + C::WVS + C::WGR); rthreads = 32 * C::WLG; }
  int lag = role;

  // number of intervals: every slot of every unit of this CTA, plus the pipeline depth
  int total = 3;
  {
    StStreamCursor c;
    for (int u = blockIdx.x; u < P.nunits; u += gridDim.x) { c.unit = u; c.decode(P, C::TW); total += c.nslots; }
  }
  StStreamCursor cur;
  cur.init(P, C::TW, blockIdx.x);
  const bool want_hr = P.ds_hr != nullptr;
  const bool norm = P.normalize != 0;
  float lsum = 0.f;

  if (role == 0 && cur.unit < P.nunits) lg_issue<C>(sStage, P, cur, rtid, rthreads);

  long long dbg_work = 0, dbg_wait = 0;
#pragma unroll 1
  for (int t = 0; t < total; ++t) {
#ifndef SRST_EMULATE
    const long long dbg_t0 = P.debug ? clock64() : 0;
#endif
    if (lag > 0) {
      --lag;
    } else if (cur.unit < P.nunits) {
      const int S = cur.S, k = cur.k;
      const size_t img_off = (size_t)cur.b * 3 * plane;
      if (role == 0) {
        // ---- LG: the RGB rows [ya-20+16k, +16) x cols [x0-12, +GW) of both images were staged during
        // the previous interval: convert them to gray (ring slot S%3), then stage the next slot.
        cp_async_wait_all();
        bar_sync(kBarStreamLG, rthreads);  // every loader thread's copies have landed
        lg_commit<C>(sG, sStage, cur, rtid, rthreads);
        bar_sync(kBarStreamLG, rthreads);  // staging fully consumed
        StStreamCursor nxt = cur;
        nxt.advance(P, C::TW, (int)gridDim.x);
        if (nxt.unit < P.nunits) lg_issue<C>(sStage, P, nxt, rtid, rthreads);
      } else if (role == 1) {
        // ---- GR: Ix, Iy of rows [ya-24+16k, +16) x cols [x0-8, +DW) -> D ring slot S%3 (k >= 1)
        if (k >= 1) {
          constexpr int SEGS = C::DW / 4;
          const int dyb = cur.ya - 24 + 16 * k;
          const int gp0 = ((S % 3) * C::RP + 3 * C::RP - 3) % (3 * C::RP);  // gray ring pair of the first input row pair
          const int dbase = (S % 3) * C::RP;
#pragma unroll 1
          for (int it = rtid; it < 2 * C::RP * SEGS; it += rthreads) {
            const int img = it / (C::RP * SEGS), rem = it - img * (C::RP * SEGS);
            const int seg = rem / C::RP, q = rem - seg * C::RP;
            const int dx0 = 4 * seg;
            const int gy = dyb + 2 * q, gx0 = cur.x0 - C::HXD + dx0;
            float2 Ix[4], Iy[4];
            if (gy + 1 >= 0 && gy < H && gx0 + 3 >= 0 && gx0 < W) {
              const float* rows[3];
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                int pr = gp0 + q + i;
                if (pr >= 3 * C::RP) pr -= 3 * C::RP;
                rows[i] = sG + img * C::G_FLOATS + pr * C::PG + 2 * (dx0 + C::BW_LO);
              }
              grad_rowpair_rows<C::RG, 4, C::BWIN, C::OFF - C::BW_LO, true>(rows, rows, tp, Ix, Iy);
              const bool r0 = gy >= 0, r1 = gy + 1 < H;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const bool okc = (gx0 + j >= 0) && (gx0 + j < W);
                Ix[j].x = (okc && r0) ? Ix[j].x : 0.f;
                Ix[j].y = (okc && r1) ? Ix[j].y : 0.f;
                Iy[j].x = (okc && r0) ? Iy[j].x : 0.f;
                Iy[j].y = (okc && r1) ? Iy[j].y : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) { Ix[j] = make_float2(0.f, 0.f); Iy[j] = make_float2(0.f, 0.f); }
            }
            float* o0 = sD + (img * 2 + 0) * C::D_FLOATS + (dbase + q) * C::PD + 2 * dx0;
            float* o1 = sD + (img * 2 + 1) * C::D_FLOATS + (dbase + q) * C::PD + 2 * dx0;
            st4(o0, make_float4(Ix[0].x, Ix[0].y, Ix[1].x, Ix[1].y));
            st4(o0 + 4, make_float4(Ix[2].x, Ix[2].y, Ix[3].x, Ix[3].y));
            st4(o1, make_float4(Iy[0].x, Iy[0].y, Iy[1].x, Iy[1].y));
            st4(o1 + 4, make_float4(Iy[2].x, Iy[2].y, Iy[3].x, Iy[3].y));
          }
        }
      } else if (role == 2) {
        // ---- VS: vertical rho-pass, rows [ya+16(k-2), +16): D ring slots (S-1)%3, S%3 -> V ring slot S%2
        if (k >= 2) {
          const int din0 = ((S + 2) % 3) * C::RP, din1 = (S % 3) * C::RP;
          const int vbase = (S % 2) * C::RP;
#pragma unroll 1
          for (int it = rtid; it < 2 * C::DW; it += rthreads) {
            const int img = it / C::DW, dx = it - img * C::DW;
            const int gx = cur.x0 - C::HXD + dx;
            float2 acc[3][C::RP];
#pragma unroll
            for (int j = 0; j < C::RP; ++j) {
              acc[0][j] = make_float2(0.f, 0.f); acc[1][j] = make_float2(0.f, 0.f); acc[2][j] = make_float2(0.f, 0.f);
            }
            if (gx >= 0 && gx < W) {
              const float* d0 = sD + (img * 2 + 0) * C::D_FLOATS + 2 * dx;
              const float* d1 = sD + (img * 2 + 1) * C::D_FLOATS + 2 * dx;
#pragma unroll
              for (int rq = 0; rq < 2 * C::RP; ++rq) {
                const int pr = (rq < C::RP) ? (din0 + rq) : (din1 + rq - C::RP);
                const float2 ix = ld2(d0 + pr * C::PD), iy = ld2(d1 + pr * C::PD);
                const float2 pxx = make_float2(ix.x * ix.x, ix.y * ix.y);
                const float2 pyy = make_float2(iy.x * iy.x, iy.y * iy.y);
                const float2 pxy = make_float2(ix.x * iy.x, ix.y * iy.y);
#pragma unroll
                for (int jp = 0; jp < C::RP; ++jp) {
                  const int u0 = 2 * rq - 2 * jp;
                  if (u0 >= 0 && u0 <= 2 * C::RK + 1) {
                    acc[0][jp] = ffma2(bcast2(pxx.x), tp.kp[u0], acc[0][jp]);
                    acc[1][jp] = ffma2(bcast2(pyy.x), tp.kp[u0], acc[1][jp]);
                    acc[2][jp] = ffma2(bcast2(pxy.x), tp.kp[u0], acc[2][jp]);
                  }
                  if (u0 + 1 >= 0 && u0 + 1 <= 2 * C::RK + 1) {
                    acc[0][jp] = ffma2(bcast2(pxx.y), tp.kp[u0 + 1], acc[0][jp]);
                    acc[1][jp] = ffma2(bcast2(pyy.y), tp.kp[u0 + 1], acc[1][jp]);
                    acc[2][jp] = ffma2(bcast2(pxy.y), tp.kp[u0 + 1], acc[2][jp]);
                  }
                }
              }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float* o = sV + (img * 3 + c) * C::V_FLOATS + vbase * C::PV + 2 * dx;
#pragma unroll
              for (int jp = 0; jp < C::RP; ++jp) st2(o + jp * C::PV, acc[c][jp]);
            }
          }
        }
      } else {
        // ---- HC: horizontal rho-pass of both images + per-pixel chain, rows [ya+16(k-2), +16)
        if (k >= 2) {
          constexpr int CG = C::TW / 2;
          const int vbase = (S % 2) * C::RP;
          const int gyb = cur.ya + 16 * (k - 2);
#pragma unroll 1
          for (int it = rtid; it < C::RP * CG; it += rthreads) {
            const int q = it / CG, cg = it - q * CG;  // lanes <-> adjacent column pairs: conflict-free LDS.128, coalesced stores
            const int ox = 2 * cg;
            const int gy0 = gyb + 2 * q, gx0 = cur.x0 + ox;
            if (gy0 >= cur.yb || gx0 >= W) continue;
            float2 S1[3][2], S2[3][2];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              smooth_h_rowpair_n<C::RK, 2, C::HWIN, C::HXD>(sV + (0 * 3 + c) * C::V_FLOATS + (vbase + q) * C::PV + 2 * ox, tp, S1[c]);
              smooth_h_rowpair_n<C::RK, 2, C::HWIN, C::HXD>(sV + (1 * 3 + c) * C::V_FLOATS + (vbase + q) * C::PV + 2 * ox, tp, S2[c]);
            }
            const bool row1 = gy0 + 1 < cur.yb;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              StPixelGrad2 G;
              G.da = G.db = G.dc = G.de = G.df = G.dh = make_float2(0.f, 0.f);
              const float2 d = want_hr ? st_pixel2<true, true>(S1[0][j], S1[1][j], S1[2][j], S2[0][j], S2[1][j], S2[2][j], norm, P.eps, G)
                                       : st_pixel2<true, false>(S1[0][j], S1[1][j], S1[2][j], S2[0][j], S2[1][j], S2[2][j], norm, P.eps, G);
              if (gx0 + j < W) {
                lsum += d.x + (row1 ? d.y : 0.f);
                const size_t o = img_off + (size_t)gy0 * W + gx0 + j;
                if (P.ds_sr) {
                  P.ds_sr[o] = G.da.x; P.ds_sr[o + plane] = G.db.x; P.ds_sr[o + 2 * plane] = G.dc.x;
                  if (row1) { P.ds_sr[o + W] = G.da.y; P.ds_sr[o + plane + W] = G.db.y; P.ds_sr[o + 2 * plane + W] = G.dc.y; }
                }
                if (want_hr) {
                  P.ds_hr[o] = G.de.x; P.ds_hr[o + plane] = G.df.x; P.ds_hr[o + 2 * plane] = G.dh.x;
                  if (row1) { P.ds_hr[o + W] = G.de.y; P.ds_hr[o + plane + W] = G.df.y; P.ds_hr[o + 2 * plane + W] = G.dh.y; }
                }
              }
            }
          }
        }
      }
      cur.advance(P, C::TW, (int)gridDim.x);
    }
#ifndef SRST_EMULATE
    const long long dbg_t1 = P.debug ? clock64() : 0;
#endif
    __syncthreads();
#ifndef SRST_EMULATE
    if (P.debug) { dbg_work += dbg_t1 - dbg_t0; dbg_wait += clock64() - dbg_t1; }
#endif
  }
  if (P.debug && blockIdx.x == 0 && (tid & 31) == 0) { P.debug[2 * warp] = dbg_work; P.debug[2 * warp + 1] = dbg_wait; }

  // Deterministic loss reduction (same scheme as the tiled kernel).
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0) s_red[warp] = lsum;
  __syncthreads();
  if (tid == 0) {
    float bs = 0.f;
    for (int w = 0; w < C::WHC; ++w) bs += s_red[w];
    P.partials[blockIdx.x] = bs;
    __threadfence();
    const unsigned int tk = atomicAdd(P.ticket, 1u);
    s_last = (tk == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (tid < 32) {
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
  }
}

}  // namespace srst
