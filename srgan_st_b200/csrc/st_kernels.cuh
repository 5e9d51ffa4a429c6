// Fused structure-tensor loss kernels for sm_100a (forward and backward).
//
// What the reference does with ~100 ATen launches per direction (loss.py:399-413,
// utils.py:212-279) is done here by ONE kernel per direction:
//
//   st_forward_kernel : RGB tile (+halo) of SR and HR -> grayscale -> Gaussian-derivative
//                       gradients Ix, Iy -> products -> separable rho-smoothing -> per-pixel
//                       det-normalise, adj(S1)*S2, eigenvalues, log-distance  -> block partial of
//                       the loss AND the per-pixel d(distance)/d(Jxx,Jyy,Jxy) ("ds" planes).
//   st_backward_kernel: ds planes -> adjoint rho-smoothing -> product rule with recomputed Ix, Iy
//                       -> adjoint derivative filters -> grayscale weights -> d_img.
//
// Both are FP32 CUDA-core stencils (no tensor cores by design: see DESIGN.md), tiled in shared
// memory with the halo recomputed per tile.  Thread mapping alternates between "a lane owns 8
// consecutive columns of one row" (horizontal filters, LDS.128 on a pitch == 4 mod 8) and "a lane
// owns one column of RS rows" (vertical filters, conflict-free LDS.32), so every filter pass is
// register-blocked: each shared-memory value feeds >= 5 FMAs.  Filter taps live in kernel
// parameters (constant bank) and every tap index is a compile-time constant after unrolling, so
// the FMAs take the tap as a constant/uniform operand.
#pragma once
#include "srst_device.cuh"

namespace srst {

template <int RG, int RK>
struct StTaps {
  float g[2 * RG + 1];   // Gaussian(sigma), utils.py:194-205
  float dg[2 * RG + 1];  // its derivative taps, utils.py:206
  float k[2 * RK + 1];   // Gaussian(rho)
};

template <int RG, int RK>
struct StFwdParams {
  const float* sr;
  const float* hr;
  float* ds_sr;  // [B,3,H,W] or null
  float* ds_hr;  // [B,3,H,W] or null
  float* partials;
  unsigned int* ticket;
  float* loss_out;
  int B, H, W, tiles_x, tiles_y;
  int normalize;
  int vec4;  // 1: W % 4 == 0 and all base pointers 16-byte aligned
  float eps;
  float inv_count;
  StTaps<RG, RK> taps;
};

template <int RG, int RK>
struct StBwdParams {
  const float* img;
  const float* ds;
  const float* grad_out;
  float* d_img;
  int B, H, W, tiles_x, tiles_y;
  int vec4;
  float inv_count;
  StTaps<RG, RK> taps;
};

// ------------------------------------------------------------------------------------------------
// Per-pixel chain: utils.py:236-279 forward and its adjoint.
//   S1 = (a,b,c) raw SR tensor (Jxx,Jyy,Jxy); S2 = (e,f,h) raw HR tensor.
// The discriminant is evaluated as (A-B)^2 + 4CD, algebraically equal to the reference's
// (A+B)^2 - 4(AB-CD) (utils.py:261) but without its catastrophic cancellation near A=B=1.
// NaN/clamp behaviour follows torch: clamp(min) keeps NaN; clamp gradient passes where x >= min.
// ------------------------------------------------------------------------------------------------
struct StPixelGrad {
  float da, db, dc;  // d dist / d (a,b,c)
  float de, df, dh;  // d dist / d (e,f,h)
};

template <bool WANT_SR, bool WANT_HR>
SRST_DEV float st_pixel(float a, float b, float c, float e, float f, float h, bool normalize, float eps,
                        StPixelGrad& G) {
  float iq1 = 1.0f, iq2 = 1.0f;
  if (normalize) {
    iq1 = 1.0f / sqrtf(a * b - c * c + eps);
    iq2 = 1.0f / sqrtf(e * f - h * h + eps);
  }
  const float ah = a * iq1, bh = b * iq1, ch = c * iq1;
  const float eh = e * iq2, fh = f * iq2, hh = h * iq2;
  const float chh = ch * hh;
  const float A = bh * eh - chh;
  const float Bm = ah * fh - chh;
  const float Cc = bh * hh - ch * fh;
  const float Dd = ah * hh - ch * eh;
  const float T = A + Bm;
  const float amb = A - Bm;
  const float disc_raw = amb * amb + 4.0f * (Cc * Dd);
  const float disc = (disc_raw < eps) ? eps : disc_raw;
  const float r = sqrtf(disc);
  const float l1r = 0.5f * (T - r), l2r = 0.5f * (T + r);
  const float l1 = (l1r < 1.0f) ? 1.0f : l1r;
  const float l2 = (l2r < 1.0f) ? 1.0f : l2r;
  const float L1 = logf(l1), L2 = logf(l2);
  const float d = sqrtf(L1 * L1 + L2 * L2 + eps);
  if (WANT_SR || WANT_HR) {
    const float inv_d = 1.0f / d;
    const float dl1 = (L1 * inv_d / l1) * ((l1r >= 1.0f) ? 1.0f : 0.0f);
    const float dl2 = (L2 * inv_d / l2) * ((l2r >= 1.0f) ? 1.0f : 0.0f);
    float dT = 0.5f * (dl1 + dl2);
    const float dr = 0.5f * (dl2 - dl1);
    const float ddisc = (dr / (2.0f * r)) * ((disc_raw >= eps) ? 1.0f : 0.0f);
    dT += 2.0f * T * ddisc;
    const float dd4 = 4.0f * ddisc;
    const float dA = dT - dd4 * Bm;
    const float dB = dT - dd4 * A;
    const float dC = dd4 * Dd;
    const float dD = dd4 * Cc;
    if (WANT_SR) {
      const float dah = dB * fh + dD * hh;
      const float dbh = dA * eh + dC * hh;
      const float dch = -(dA + dB) * hh - dC * fh - dD * eh;
      if (normalize) {
        // S^ = S/q, q = sqrt(det+eps): dS = dS^/q - S * <S,dS^>/(2 q^3) * d(det)/dS
        const float s = a * dah + b * dbh + c * dch;
        const float ddet = -0.5f * s * iq1 * iq1 * iq1;
        G.da = dah * iq1 + ddet * b;
        G.db = dbh * iq1 + ddet * a;
        G.dc = dch * iq1 - 2.0f * ddet * c;
      } else {
        G.da = dah; G.db = dbh; G.dc = dch;
      }
    }
    if (WANT_HR) {
      const float deh = dA * bh - dD * ch;
      const float dfh = dB * ah - dC * ch;
      const float dhh = -(dA + dB) * ch + dC * bh + dD * ah;
      if (normalize) {
        const float s = e * deh + f * dfh + h * dhh;
        const float ddet = -0.5f * s * iq2 * iq2 * iq2;
        G.de = deh * iq2 + ddet * f;
        G.df = dfh * iq2 + ddet * e;
        G.dh = dhh * iq2 - 2.0f * ddet * h;
      } else {
        G.de = deh; G.df = dfh; G.dh = dhh;
      }
    }
  }
  return d;
}

// ------------------------------------------------------------------------------------------------
// Shared building blocks
// ------------------------------------------------------------------------------------------------

// Load the RGB pixels of rows [gy0, gy0+ROWS) x cols [gx0, gx0+COLS) of image `base` ([3,H,W]),
// convert to grayscale and store into sG (pitch PITCH); zero outside the image (the reference
// zero-pads: padding='same', utils.py:219-222).  gx0 and COLS are multiples of 4.
template <int ROWS, int COLS, int PITCH, int NT>
SRST_DEV void load_gray_tile(float* sG, const float* __restrict__ base, int H, int W, int gy0, int gx0,
                             bool vec4, int tid) {
  constexpr int C4 = COLS / 4;
  const size_t plane = (size_t)H * W;
  for (int it = tid; it < ROWS * C4; it += NT) {
    const int r = it / C4, c4 = it - r * C4;
    const int gy = gy0 + r, gx = gx0 + 4 * c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H) {
      const float* p = base + (size_t)gy * W + gx;
      if (vec4) {
        if (gx >= 0 && gx < W) {  // W % 4 == 0: the group is all-in or all-out
          const float4 R = ldg4(p), Gc = ldg4(p + plane), Bc = ldg4(p + 2 * plane);
          v.x = gray_of(R.x, Gc.x, Bc.x);
          v.y = gray_of(R.y, Gc.y, Bc.y);
          v.z = gray_of(R.z, Gc.z, Bc.z);
          v.w = gray_of(R.w, Gc.w, Bc.w);
        }
      } else {
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = gx + j;
          t[j] = (x >= 0 && x < W) ? gray_of(__ldg(p + j), __ldg(p + j + plane), __ldg(p + j + 2 * plane)) : 0.f;
        }
        v = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    st4(sG + r * PITCH + 4 * c4, v);
  }
}

// Gaussian-derivative gradients for 8 consecutive pixels of one row, from a gray tile in smem.
//   Ix = (im * dg|) * g-   (derivative along H),  Iy = (im * g|) * dg-   (utils.py:219-222)
// `row0` points at gray[(first needed row)][window start]; the window is WIN floats wide and the
// 8 outputs are centred at window index CEN..CEN+7.
template <int RG, int WIN, int CEN, int PITCH, class Taps>
SRST_DEV void grad8(const float* row0, const Taps& tp, float (&Ix)[8], float (&Iy)[8]) {
  float tA[WIN], tB[WIN];
#pragma unroll
  for (int j = 0; j < WIN; ++j) { tA[j] = 0.f; tB[j] = 0.f; }
#pragma unroll
  for (int i = 0; i <= 2 * RG; ++i) {
    float v[WIN];
#pragma unroll
    for (int q = 0; q < WIN / 4; ++q) {
      const float4 t = ld4(row0 + i * PITCH + 4 * q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int j = CEN - RG; j < CEN + 8 + RG; ++j) {
      if (i != RG) tA[j] = fmaf(tp.dg[i], v[j], tA[j]);  // dg[RG] = phi*(-0)/sigma^2 == 0
      tB[j] = fmaf(tp.g[i], v[j], tB[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int t = 0; t <= 2 * RG; ++t) {
      sx = fmaf(tp.g[t], tA[CEN + j + t - RG], sx);
      if (t != RG) sy = fmaf(tp.dg[t], tB[CEN + j + t - RG], sy);
    }
    Ix[j] = sx;
    Iy[j] = sy;
  }
}

// ------------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------------
template <int TH_, int TW_, int RS_, int RG_, int RK_, int MINB_>
struct StFwdCfg {
  static constexpr int TH = TH_, TW = TW_, RS = RS_, RG = RG_, RK = RK_, MINB = MINB_;
  static constexpr int CS = 8;
  static constexpr int NT = TH * TW / CS;    // one horizontal-pass item per thread
  static constexpr int HXD = round_up4(RK);  // x halo of the gradient (D) and V regions
  static constexpr int OFF = round_up4(RG);
  static constexpr int HXG = HXD + OFF;      // x halo of the gray (G) region
  static constexpr int GH = TH + 2 * (RG + RK), GW = TW + 2 * HXG, PG = smem_pitch(GW);
  static constexpr int DH = TH + 2 * RK, DW = TW + 2 * HXD, PD = smem_pitch(DW);
  static constexpr int PV = PD;
  static constexpr int NSEG = TH / RS;
  static constexpr int BW_LO = round_dn4(OFF - RG), BW_HI = round_up4(OFF + CS + RG), BWIN = BW_HI - BW_LO;
  static constexpr int DW_LO = round_dn4(HXD - RK), DW_HI = round_up4(HXD + CS + RK), DWIN = DW_HI - DW_LO;
  static constexpr int SMEM_FLOATS = 2 * DH * PD + cmax(3 * TH * PV, GH * PG);
  static constexpr size_t SMEM_BYTES = sizeof(float) * SMEM_FLOATS;
  static_assert(TH % RS == 0 && TW % CS == 0 && NT % 32 == 0 && NT <= 1024, "bad forward tile");
};

// Phases A-D for one image of the pair: leaves the smoothed tensor (Jxx,Jyy,Jxy) of this thread's
// 8 pixels (row oy, cols ox0..ox0+7 of the tile) in S[3][8].
template <class C, class Taps>
SRST_DEV void st_tile_tensor(float* smem, const float* __restrict__ base, int H, int W, int y0, int x0,
                             bool vec4, const Taps& tp, int tid, float (&S)[3][8]) {
  float* sD0 = smem;
  float* sD1 = sD0 + C::DH * C::PD;
  float* sV = sD1 + C::DH * C::PD;
  float* sG = sV;  // gray tile aliases V: it is dead once phase B is done

  // Phase A: global -> gray tile (with halo RG+RK rows, HXG cols)
  load_gray_tile<C::GH, C::GW, C::PG, C::NT>(sG, base, H, W, y0 - (C::RG + C::RK), x0 - C::HXG, vec4, tid);
  __syncthreads();

  // Phase B: Ix, Iy on the D region; forced to zero outside the image because the reference
  // zero-pads the *products* for the rho-smoothing (utils.py:225-230).
  for (int it = tid; it < C::DH * (C::DW / 8); it += C::NT) {
    const int seg = it / C::DH, r = it - seg * C::DH;
    const int dx0 = 8 * seg;
    const int gy = y0 - C::RK + r, gx0 = x0 - C::HXD + dx0;
    float Ix[8], Iy[8];
    if (gy >= 0 && gy < H && gx0 + 7 >= 0 && gx0 < W) {
      grad8<C::RG, C::BWIN, C::OFF - C::BW_LO, C::PG>(sG + r * C::PG + dx0 + C::BW_LO, tp, Ix, Iy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = (gx0 + j >= 0) && (gx0 + j < W);
        Ix[j] = ok ? Ix[j] : 0.f;
        Iy[j] = ok ? Iy[j] : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { Ix[j] = 0.f; Iy[j] = 0.f; }
    }
    float* o0 = sD0 + r * C::PD + dx0;
    float* o1 = sD1 + r * C::PD + dx0;
    st4(o0, make_float4(Ix[0], Ix[1], Ix[2], Ix[3]));
    st4(o0 + 4, make_float4(Ix[4], Ix[5], Ix[6], Ix[7]));
    st4(o1, make_float4(Iy[0], Iy[1], Iy[2], Iy[3]));
    st4(o1 + 4, make_float4(Iy[4], Iy[5], Iy[6], Iy[7]));
  }
  __syncthreads();

  // Phase C: vertical rho-pass of the three products; a lane owns one column and RS output rows.
  for (int it = tid; it < C::DW * C::NSEG; it += C::NT) {
    const int seg = it / C::DW, dx = it - seg * C::DW;
    const int gx = x0 - C::HXD + dx;
    float acc[3][C::RS];
#pragma unroll
    for (int j = 0; j < C::RS; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; acc[2][j] = 0.f; }
    if (gx >= 0 && gx < W) {
      const float* p0 = sD0 + (seg * C::RS) * C::PD + dx;
      const float* p1 = sD1 + (seg * C::RS) * C::PD + dx;
#pragma unroll
      for (int r = 0; r < C::RS + 2 * C::RK; ++r) {
        const float ix = p0[r * C::PD], iy = p1[r * C::PD];
        const float pxx = ix * ix, pyy = iy * iy, pxy = ix * iy;
#pragma unroll
        for (int j = 0; j < C::RS; ++j) {
          const int t = r - j;
          if (t >= 0 && t <= 2 * C::RK) {
            acc[0][j] = fmaf(tp.k[t], pxx, acc[0][j]);
            acc[1][j] = fmaf(tp.k[t], pyy, acc[1][j]);
            acc[2][j] = fmaf(tp.k[t], pxy, acc[2][j]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float* o = sV + c * (C::TH * C::PV) + (seg * C::RS) * C::PV + dx;
#pragma unroll
      for (int j = 0; j < C::RS; ++j) o[j * C::PV] = acc[c][j];
    }
  }
  __syncthreads();

  // Phase D: horizontal rho-pass; a lane owns 8 consecutive columns of one row.
  {
    const int seg = tid / C::TH, oy = tid - seg * C::TH;
    const int ox0 = 8 * seg;
    constexpr int CEN = C::HXD - C::DW_LO;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* row = sV + c * (C::TH * C::PV) + oy * C::PV + ox0 + C::DW_LO;
      float v[C::DWIN];
#pragma unroll
      for (int q = 0; q < C::DWIN / 4; ++q) {
        const float4 t = ld4(row + 4 * q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t <= 2 * C::RK; ++t) s = fmaf(tp.k[t], v[CEN + j + t - C::RK], s);
        S[c][j] = s;
      }
    }
  }
  __syncthreads();  // the next image's phase A overwrites sG (== sV)
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
st_forward_kernel(const __grid_constant__ StFwdParams<C::RG, C::RK> P) {
  SRST_DYN_SMEM(float, smem);
  __shared__ float s_red[32];
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  const int tx = tile % P.tiles_x;
  tile /= P.tiles_x;
  const int ty = tile % P.tiles_y;
  const int b = tile / P.tiles_y;
  const int y0 = ty * C::TH, x0 = tx * C::TW;
  const size_t img_off = (size_t)b * 3 * P.H * P.W;

  float S1[3][8], S2[3][8];
  st_tile_tensor<C>(smem, P.sr + img_off, P.H, P.W, y0, x0, P.vec4 != 0, P.taps, tid, S1);
  st_tile_tensor<C>(smem, P.hr + img_off, P.H, P.W, y0, x0, P.vec4 != 0, P.taps, tid, S2);

  // Per-pixel chain on this thread's 8 pixels.
  const int seg = tid / C::TH, oy = tid - seg * C::TH;
  const int gy = y0 + oy, gx0 = x0 + 8 * seg;
  const bool want_sr = P.ds_sr != nullptr, want_hr = P.ds_hr != nullptr;
  const bool norm = P.normalize != 0;
  float lsum = 0.f;
  float g_sr[3][8], g_hr[3][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    StPixelGrad G;
    G.da = G.db = G.dc = G.de = G.df = G.dh = 0.f;
    float d;
    if (want_hr)
      d = st_pixel<true, true>(S1[0][j], S1[1][j], S1[2][j], S2[0][j], S2[1][j], S2[2][j], norm, P.eps, G);
    else if (want_sr)
      d = st_pixel<true, false>(S1[0][j], S1[1][j], S1[2][j], S2[0][j], S2[1][j], S2[2][j], norm, P.eps, G);
    else
      d = st_pixel<false, false>(S1[0][j], S1[1][j], S1[2][j], S2[0][j], S2[1][j], S2[2][j], norm, P.eps, G);
    const bool ok = (gy < P.H) && (gx0 + j < P.W);
    lsum += ok ? d : 0.f;
    g_sr[0][j] = G.da; g_sr[1][j] = G.db; g_sr[2][j] = G.dc;
    g_hr[0][j] = G.de; g_hr[1][j] = G.df; g_hr[2][j] = G.dh;
  }
  if (gy < P.H && gx0 < P.W) {
    const size_t plane = (size_t)P.H * P.W;
    const size_t o = img_off + (size_t)gy * P.W + gx0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (P.vec4) {
        if (want_sr) {
          st4(P.ds_sr + o + c * plane, make_float4(g_sr[c][0], g_sr[c][1], g_sr[c][2], g_sr[c][3]));
          if (gx0 + 4 < P.W) st4(P.ds_sr + o + c * plane + 4, make_float4(g_sr[c][4], g_sr[c][5], g_sr[c][6], g_sr[c][7]));
        }
        if (want_hr) {
          st4(P.ds_hr + o + c * plane, make_float4(g_hr[c][0], g_hr[c][1], g_hr[c][2], g_hr[c][3]));
          if (gx0 + 4 < P.W) st4(P.ds_hr + o + c * plane + 4, make_float4(g_hr[c][4], g_hr[c][5], g_hr[c][6], g_hr[c][7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (gx0 + j < P.W) {
            if (want_sr) P.ds_sr[o + c * plane + j] = g_sr[c][j];
            if (want_hr) P.ds_hr[o + c * plane + j] = g_hr[c][j];
          }
        }
      }
    }
  }

  // Deterministic loss reduction: block partial -> workspace; the last block to finish sums all
  // partials in a fixed order (double) and re-zeroes the workspace for the next call.
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0) s_red[tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) {
    float bs = 0.f;
    for (int w = 0; w < C::NT / 32; ++w) bs += s_red[w];
    P.partials[blockIdx.x] = bs;
    __threadfence();
    const unsigned int t = atomicAdd(P.ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
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

// ------------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------------
template <int TH_, int TW_, int RS_, int NT_, int RG_, int RK_, int MINB_>
struct StBwdCfg {
  static constexpr int TH = TH_, TW = TW_, RS = RS_, NT = NT_, RG = RG_, RK = RK_, MINB = MINB_;
  static constexpr int HXE = round_up4(RG);  // x halo of the E region (where dIx, dIy are needed)
  static constexpr int HXK = round_up4(RK);
  static constexpr int EH = TH + 2 * RG, EW = TW + 2 * HXE, PE = smem_pitch(EW);
  static constexpr int GH = EH + 2 * RG, GW = EW + 2 * HXE, PG = smem_pitch(GW);  // gray region
  static constexpr int VW = EW + 2 * HXK, PV = smem_pitch(VW);                     // vertical-pass output
  static constexpr int NSEG = EH / RS;
  // gradient window (same geometry as the forward phase B, output cols relative to the E region)
  static constexpr int BW_LO = round_dn4(HXE - RG), BW_HI = round_up4(HXE + 8 + RG), BWIN = BW_HI - BW_LO;
  // horizontal rho-pass window on the V region
  static constexpr int DW_LO = round_dn4(HXK - RK), DW_HI = round_up4(HXK + 8 + RK), DWIN = DW_HI - DW_LO;
  // smem: V (3 planes) | I (Ix, Iy) | dI (dIx, dIy);  gray aliases dI (dead before dI is written)
  static constexpr int V_FLOATS = 3 * EH * PV, I_FLOATS = 2 * EH * PE;
  static constexpr int X_FLOATS = cmax(2 * EH * PE, GH * PG);
  static constexpr int SMEM_FLOATS = V_FLOATS + I_FLOATS + X_FLOATS;
  static constexpr size_t SMEM_BYTES = sizeof(float) * SMEM_FLOATS;
  static_assert(EH % RS == 0 && TW % 8 == 0 && NT % 32 == 0 && NT <= 1024, "bad backward tile");
};

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
st_backward_kernel(const __grid_constant__ StBwdParams<C::RG, C::RK> P) {
  SRST_DYN_SMEM(float, smem);
  float* sV = smem;                  // [3][EH][PV]  vertical rho-pass of ds
  float* sI0 = sV + C::V_FLOATS;     // Ix [EH][PE]
  float* sI1 = sI0 + C::EH * C::PE;  // Iy
  float* sX = sI1 + C::EH * C::PE;   // gray [GH][PG], later dIx|dIy [2][EH][PE]
  float* sG = sX;
  float* sdI0 = sX;
  float* sdI1 = sX + C::EH * C::PE;

  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  const int tx = tile % P.tiles_x;
  tile /= P.tiles_x;
  const int ty = tile % P.tiles_y;
  const int b = tile / P.tiles_y;
  const int y0 = ty * C::TH, x0 = tx * C::TW;
  const int H = P.H, W = P.W;
  const size_t plane = (size_t)H * W;
  const size_t img_off = (size_t)b * 3 * plane;
  const auto& tp = P.taps;

  // Phase A': gray tile of the image (halo 2*RG rows, 2*HXE cols)
  load_gray_tile<C::GH, C::GW, C::PG, C::NT>(sG, P.img + img_off, H, W, y0 - 2 * C::RG, x0 - 2 * C::HXE,
                                             P.vec4 != 0, tid);

  // Phase C': vertical rho-pass of the three ds planes straight from global memory (coalesced:
  // consecutive lanes read consecutive columns).  The adjoint of the zero-padded symmetric
  // smoothing is the same zero-padded smoothing.
  {
    const float* dsb = P.ds + img_off;
    for (int it = tid; it < C::VW * C::NSEG; it += C::NT) {
      const int seg = it / C::VW, vx = it - seg * C::VW;
      const int gx = x0 - C::HXE - C::HXK + vx;
      const int gyb = y0 - C::RG - C::RK + seg * C::RS;  // first input row
      const bool colok = gx >= 0 && gx < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float acc[C::RS];
#pragma unroll
        for (int j = 0; j < C::RS; ++j) acc[j] = 0.f;
        if (colok) {
          const float* p = dsb + c * plane + gx;
          float v[C::RS + 2 * C::RK];
#pragma unroll
          for (int r = 0; r < C::RS + 2 * C::RK; ++r) {
            const int gy = gyb + r;
            v[r] = (gy >= 0 && gy < H) ? __ldg(p + (size_t)gy * W) : 0.f;
          }
#pragma unroll
          for (int r = 0; r < C::RS + 2 * C::RK; ++r) {
#pragma unroll
            for (int j = 0; j < C::RS; ++j) {
              const int t = r - j;
              if (t >= 0 && t <= 2 * C::RK) acc[j] = fmaf(tp.k[t], v[r], acc[j]);
            }
          }
        }
        float* o = sV + c * (C::EH * C::PV) + (seg * C::RS) * C::PV + vx;
#pragma unroll
        for (int j = 0; j < C::RS; ++j) o[j * C::PV] = acc[j];
      }
    }
  }
  __syncthreads();

  // Phase B': recompute Ix, Iy on the E region (zero outside the image).
  for (int it = tid; it < C::EH * (C::EW / 8); it += C::NT) {
    const int seg = it / C::EH, r = it - seg * C::EH;
    const int ex0 = 8 * seg;
    const int gy = y0 - C::RG + r, gx0 = x0 - C::HXE + ex0;
    float Ix[8], Iy[8];
    if (gy >= 0 && gy < H && gx0 + 7 >= 0 && gx0 < W) {
      grad8<C::RG, C::BWIN, C::HXE - C::BW_LO, C::PG>(sG + r * C::PG + ex0 + C::BW_LO, tp, Ix, Iy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = (gx0 + j >= 0) && (gx0 + j < W);
        Ix[j] = ok ? Ix[j] : 0.f;
        Iy[j] = ok ? Iy[j] : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { Ix[j] = 0.f; Iy[j] = 0.f; }
    }
    float* o0 = sI0 + r * C::PE + ex0;
    float* o1 = sI1 + r * C::PE + ex0;
    st4(o0, make_float4(Ix[0], Ix[1], Ix[2], Ix[3]));
    st4(o0 + 4, make_float4(Ix[4], Ix[5], Ix[6], Ix[7]));
    st4(o1, make_float4(Iy[0], Iy[1], Iy[2], Iy[3]));
    st4(o1 + 4, make_float4(Iy[4], Iy[5], Iy[6], Iy[7]));
  }
  __syncthreads();  // gray is dead from here on: sdI may overwrite it

  // Phase D': horizontal rho-pass -> E = K*ds at the E-region pixels, then the product rule
  //   dIx = 2 Ix Exx + Iy Exy ,  dIy = 2 Iy Eyy + Ix Exy      (adjoint of utils.py:225-229)
  for (int it = tid; it < C::EH * (C::EW / 8); it += C::NT) {
    const int seg = it / C::EH, r = it - seg * C::EH;
    const int ex0 = 8 * seg;
    constexpr int CEN = C::HXK - C::DW_LO;
    float E[3][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* row = sV + c * (C::EH * C::PV) + r * C::PV + ex0 + C::DW_LO;
      float v[C::DWIN];
#pragma unroll
      for (int q = 0; q < C::DWIN / 4; ++q) {
        const float4 t = ld4(row + 4 * q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t <= 2 * C::RK; ++t) s = fmaf(tp.k[t], v[CEN + j + t - C::RK], s);
        E[c][j] = s;
      }
    }
    const float4 ixa = ld4(sI0 + r * C::PE + ex0), ixb = ld4(sI0 + r * C::PE + ex0 + 4);
    const float4 iya = ld4(sI1 + r * C::PE + ex0), iyb = ld4(sI1 + r * C::PE + ex0 + 4);
    const float ix[8] = {ixa.x, ixa.y, ixa.z, ixa.w, ixb.x, ixb.y, ixb.z, ixb.w};
    const float iy[8] = {iya.x, iya.y, iya.z, iya.w, iyb.x, iyb.y, iyb.z, iyb.w};
    float dx[8], dy[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dx[j] = 2.0f * ix[j] * E[0][j] + iy[j] * E[2][j];
      dy[j] = 2.0f * iy[j] * E[1][j] + ix[j] * E[2][j];
    }
    float* o0 = sdI0 + r * C::PE + ex0;
    float* o1 = sdI1 + r * C::PE + ex0;
    st4(o0, make_float4(dx[0], dx[1], dx[2], dx[3]));
    st4(o0 + 4, make_float4(dx[4], dx[5], dx[6], dx[7]));
    st4(o1, make_float4(dy[0], dy[1], dy[2], dy[3]));
    st4(o1 + 4, make_float4(dy[4], dy[5], dy[6], dy[7]));
  }
  __syncthreads();

  // Phase E': adjoint of the derivative filters.  The adjoint of a zero-padded correlation is the
  // correlation with flipped taps; g is symmetric and dg antisymmetric, so
  //   dgray = -[ (dIx * dg|) * g-  +  (dIy * g|) * dg- ]
  // i.e. the forward gradient operators applied to dIx and dIy, negated.
  const float scale = __ldg(P.grad_out) * P.inv_count;
  for (int it = tid; it < C::TH * (C::TW / 8); it += C::NT) {
    const int seg = it / C::TH, oy = it - seg * C::TH;
    const int ox0 = 8 * seg;
    const int gy = y0 + oy, gx0 = x0 + ox0;
    if (gy >= H || gx0 >= W) continue;
    constexpr int WIN = C::BWIN, CEN = C::HXE - C::BW_LO;
    float tA[WIN], tB[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) { tA[j] = 0.f; tB[j] = 0.f; }
#pragma unroll
    for (int i = 0; i <= 2 * C::RG; ++i) {
      const float* ra = sdI0 + (oy + i) * C::PE + ox0 + C::BW_LO;
      const float* rb = sdI1 + (oy + i) * C::PE + ox0 + C::BW_LO;
      float u[WIN], v[WIN];
#pragma unroll
      for (int q = 0; q < WIN / 4; ++q) {
        const float4 t = ld4(ra + 4 * q);
        u[4 * q] = t.x; u[4 * q + 1] = t.y; u[4 * q + 2] = t.z; u[4 * q + 3] = t.w;
        const float4 s = ld4(rb + 4 * q);
        v[4 * q] = s.x; v[4 * q + 1] = s.y; v[4 * q + 2] = s.z; v[4 * q + 3] = s.w;
      }
#pragma unroll
      for (int j = CEN - C::RG; j < CEN + 8 + C::RG; ++j) {
        if (i != C::RG) tA[j] = fmaf(tp.dg[i], u[j], tA[j]);
        tB[j] = fmaf(tp.g[i], v[j], tB[j]);
      }
    }
    float dgr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t <= 2 * C::RG; ++t) {
        s = fmaf(tp.g[t], tA[CEN + j + t - C::RG], s);
        if (t != C::RG) s = fmaf(tp.dg[t], tB[CEN + j + t - C::RG], s);
      }
      dgr[j] = -s * scale;
    }
    float* o = P.d_img + img_off + (size_t)gy * W + gx0;
    const float coef[3] = {kGrayR, kGrayG, kGrayB};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (P.vec4) {
        st4(o + c * plane, make_float4(coef[c] * dgr[0], coef[c] * dgr[1], coef[c] * dgr[2], coef[c] * dgr[3]));
        if (gx0 + 4 < W)
          st4(o + c * plane + 4, make_float4(coef[c] * dgr[4], coef[c] * dgr[5], coef[c] * dgr[6], coef[c] * dgr[7]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (gx0 + j < W) o[c * plane + j] = coef[c] * dgr[j];
      }
    }
  }
}

}  // namespace srst
