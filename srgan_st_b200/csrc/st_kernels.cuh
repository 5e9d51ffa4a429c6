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
// Both are FP32 CUDA-core stencils (no tensor cores by design: see DESIGN.md).  The measured
// bound on B200 is the FP32 pipe (128 lane-FMA/clk/SM, tools/ubench_fma.cu), not HBM, so the
// design minimises issue slots per FMA:
//   * every shared-memory plane is stored ROW-PAIR INTERLEAVED: element (row r, col c) lives at
//     (r/2)*pitch + 2*c + (r&1), so a float2 holds one column of two adjacent rows;
//   * every filter pass works on such row pairs with packed fma.rn.f32x2 (SASS FFMA2): horizontal
//     passes multiply a register pair by one tap broadcast from the constant bank, vertical passes
//     multiply one value (broadcast) by the tap PAIR (k[t], k[t-1]) that maps it onto the two
//     output rows.  Both operand forms are native (FFMA2 Rd, Ra.F32, URb.F32x2, Rc), so a 17-tap
//     pass costs 17 issue slots per pixel pair and no register moves;
//   * horizontal passes are register-blocked (4 columns x 2 rows per thread, LDS.128 on a pitch
//     == 4 mod 8), vertical passes own one column and RS rows (conflict-free LDS.64/STS.64);
//   * taps live in kernel parameters and every tap index is a compile-time constant.
#pragma once
#include "srst_device.cuh"

namespace srst {

template <int RG, int RK>
struct StTaps {
  float g[2 * RG + 1];    // Gaussian(sigma), utils.py:194-205
  float dg[2 * RG + 1];   // its derivative taps, utils.py:206
  float k[2 * RK + 1];    // Gaussian(rho)
  // tap pairs (t[u], t[u-1]) with zeros outside the support: input row u of a row pair's window
  // contributes t[u] to the even output row and t[u-1] to the odd one.
  float2 gp[2 * RG + 2];
  float2 dgp[2 * RG + 2];
  float2 kp[2 * RK + 2];
};

template <int RG, int RK>
struct StFwdParams {
  const float* sr;
  const float* hr;
  float* ds_sr;  // [B,3,H,W] or null
  float* ds_hr;  // [B,3,H,W] or null
  float* gray_sr;  // [B,H,W] or null: grayscale planes saved for the backward pass
  float* gray_hr;
  float* partials;
  float* px_partials;  // null, or per-CTA partials of sum (sr-hr)^2: the fused "Pixel" MSE term (warmup.py:88-96)
  unsigned int* ticket;
  float* loss_out;     // [0] = ST loss; [1] = MSE when px_partials is set
  int B, H, W, tiles_x, tiles_y;
  int normalize;
  int vec4;  // 1: W % 4 == 0 and all base pointers 16-byte aligned
  float eps;
  float inv_count;
  long long* debug;  // optional (SRST_ST_DEBUG=1): per-warp phase time stamps of CTA 0
  StTaps<RG, RK> taps;
};

template <int RG, int RK>
struct StBwdParams {
  SrstTmap ds_map;    // tensor map of ds viewed as [B*3][H][W] (valid when use_tma)
  SrstTmap gray_map;  // tensor map of the saved gray planes [B][H][W] (valid when use_gray)
  const float* img;
  const float* ds;
  const float* grad_out;
  const float* px_other;  // null, or the other image of the pair: adds grad_px * 2 (img - other) / (3 B H W) to d_img
  const float* grad_px;   // upstream gradient of the MSE term (device scalar)
  float* d_img;
  int B, H, W, tiles_x, tiles_y;
  int vec4;
  int use_tma;
  int use_gray;  // 1: the gray tile comes from gray_map by TMA instead of RGB loads + conversion
  float inv_count;
  int early_ctas;  // CTAs [0, early_ctas) -- the first wave -- load the image and rebuild Ix, Iy BEFORE waiting for the
                   // previous kernel of the stream (programmatic dependent launch): that work only reads inputs
  long long* debug;
  StTaps<RG, RK> taps;
};

SRST_DEV float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
SRST_DEV float2 bcast2(float v) { return make_float2(v, v); }
SRST_DEV float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
SRST_DEV void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }

// ------------------------------------------------------------------------------------------------
// Per-pixel chain: utils.py:236-279 forward and its adjoint.
//   S1 = (a,b,c) raw SR tensor (Jxx,Jyy,Jxy); S2 = (e,f,h) raw HR tensor.
// The discriminant is evaluated as (A-B)^2 + 4CD, algebraically equal to the reference's
// (A+B)^2 - 4(AB-CD) (utils.py:261) but without its catastrophic cancellation near A=B=1.
// NaN/clamp behaviour follows torch: clamp(min) keeps NaN; clamp gradient passes where x >= min.
// Transcendentals use the SFU (MUFU.RSQ/LG2/RCP): rsqrt of the two determinants gets one Newton
// step (they scale everything downstream); the rest stay at SFU accuracy (<= 2 ulp), far inside
// the 1e-5 / 1e-4 parity budget, at ~1/10 of the issue slots of sqrtf/logf/division.
// ------------------------------------------------------------------------------------------------
struct StPixelGrad {
  float da, db, dc;  // d dist / d (a,b,c)
  float de, df, dh;  // d dist / d (e,f,h)
};

SRST_DEV float rsqrt_nr(float x) {
  const float y = fast_rsqrt(x);
  return y * fmaf(-0.5f * x, y * y, 1.5f);  // one Newton step; NaN for x < 0 like 1/sqrt(x)
}

template <bool WANT_SR, bool WANT_HR>
SRST_DEV float st_pixel(float a, float b, float c, float e, float f, float h, bool normalize, float eps,
                        StPixelGrad& G) {
  constexpr float kLn2 = 0.6931471805599453f;
  float iq1 = 1.0f, iq2 = 1.0f;
  if (normalize) {
    iq1 = rsqrt_nr(fmaf(a, b, -c * c) + eps);
    iq2 = rsqrt_nr(fmaf(e, f, -h * h) + eps);
  }
  const float ah = a * iq1, bh = b * iq1, ch = c * iq1;
  const float eh = e * iq2, fh = f * iq2, hh = h * iq2;
  const float chh = ch * hh;
  const float A = fmaf(bh, eh, -chh);
  const float Bm = fmaf(ah, fh, -chh);
  const float Cc = fmaf(bh, hh, -ch * fh);
  const float Dd = fmaf(ah, hh, -ch * eh);
  const float T = A + Bm;
  const float amb = A - Bm;
  const float disc_raw = fmaf(amb, amb, 4.0f * (Cc * Dd));
  const float disc = (disc_raw < eps) ? eps : disc_raw;
  const float ir = fast_rsqrt(disc);
  const float r = disc * ir;
  const float hT = 0.5f * T;
  const float l1r = fmaf(-0.5f, r, hT), l2r = fmaf(0.5f, r, hT);
  const float l1 = (l1r < 1.0f) ? 1.0f : l1r;
  const float l2 = (l2r < 1.0f) ? 1.0f : l2r;
  // natural logs: det M == 1 makes l1 <= 1 <= l2, so L1 is 0 except for rounding stragglers
  const float L1 = kLn2 * fast_lg2(l1), L2 = kLn2 * fast_lg2(l2);
  const float arg = fmaf(L1, L1, fmaf(L2, L2, eps));
  const float inv_d = fast_rsqrt(arg);
  const float d = arg * inv_d;
  if (WANT_SR || WANT_HR) {
    const float dl1 = (l1r >= 1.0f) ? (L1 * inv_d) * fast_rcp(l1) : (l1r * 0.0f);  // x*0 keeps NaN
    const float dl2 = (l2r >= 1.0f) ? (L2 * inv_d) * fast_rcp(l2) : (l2r * 0.0f);
    const float dr = 0.5f * (dl2 - dl1);
    const float ddisc = (disc_raw >= eps) ? (0.5f * dr * ir) : (disc_raw * 0.0f);
    const float dT = fmaf(2.0f * T, ddisc, 0.5f * (dl1 + dl2));
    const float dd4 = 4.0f * ddisc;
    const float dA = fmaf(-dd4, Bm, dT);
    const float dB = fmaf(-dd4, A, dT);
    const float dC = dd4 * Dd;
    const float dD = dd4 * Cc;
    if (WANT_SR) {
      const float dah = fmaf(dB, fh, dD * hh);
      const float dbh = fmaf(dA, eh, dC * hh);
      const float dch = -fmaf(dA + dB, hh, fmaf(dC, fh, dD * eh));
      if (normalize) {
        // S^ = S/q, q = sqrt(det+eps): dS = dS^/q - S * <S,dS^>/(2 q^3) * d(det)/dS
        const float s = fmaf(a, dah, fmaf(b, dbh, c * dch));
        const float ddet = (-0.5f * s) * (iq1 * iq1) * iq1;
        G.da = fmaf(dah, iq1, ddet * b);
        G.db = fmaf(dbh, iq1, ddet * a);
        G.dc = fmaf(dch, iq1, -2.0f * ddet * c);
      } else {
        G.da = dah; G.db = dbh; G.dc = dch;
      }
    }
    if (WANT_HR) {
      const float deh = fmaf(dA, bh, -dD * ch);
      const float dfh = fmaf(dB, ah, -dC * ch);
      const float dhh = fmaf(-(dA + dB), ch, fmaf(dC, bh, dD * ah));
      if (normalize) {
        const float s = fmaf(e, deh, fmaf(f, dfh, h * dhh));
        const float ddet = (-0.5f * s) * (iq2 * iq2) * iq2;
        G.de = fmaf(deh, iq2, ddet * f);
        G.df = fmaf(dfh, iq2, ddet * e);
        G.dh = fmaf(dhh, iq2, -2.0f * ddet * h);
      } else {
        G.de = deh; G.df = dfh; G.dh = dhh;
      }
    }
  }
  return d;
}

// Packed form of st_pixel: the same chain on TWO pixels at once (.x = even row, .y = odd row of a
// row pair), every add/mul/fma issued as FADD2/FMUL2/FFMA2; SFU ops and selects stay per lane.
struct StPixelGrad2 {
  float2 da, db, dc, de, df, dh;
};
SRST_DEV float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
SRST_DEV float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
SRST_DEV float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
SRST_DEV float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); }
SRST_DEV float2 rsq2(float2 a) { return make_float2(fast_rsqrt(a.x), fast_rsqrt(a.y)); }
SRST_DEV float2 rcp2(float2 a) { return make_float2(fast_rcp(a.x), fast_rcp(a.y)); }
SRST_DEV float2 lg22(float2 a) { return make_float2(fast_lg2(a.x), fast_lg2(a.y)); }
SRST_DEV float2 rsqrt_nr2(float2 x) {
  const float2 y = rsq2(x);
  return mul2(y, ffma2(mul2(bcast2(-0.5f), x), mul2(y, y), bcast2(1.5f)));
}

template <bool WANT_SR, bool WANT_HR>
SRST_DEV float2 st_pixel2(float2 a, float2 b, float2 c, float2 e, float2 f, float2 h, bool normalize, float eps,
                          StPixelGrad2& G) {
  constexpr float kLn2 = 0.6931471805599453f;
  const float2 eps2 = bcast2(eps);
  float2 iq1 = bcast2(1.0f), iq2 = bcast2(1.0f);
  if (normalize) {
    iq1 = rsqrt_nr2(add2(ffma2(a, b, neg2(mul2(c, c))), eps2));
    iq2 = rsqrt_nr2(add2(ffma2(e, f, neg2(mul2(h, h))), eps2));
  }
  const float2 ah = mul2(a, iq1), bh = mul2(b, iq1), ch = mul2(c, iq1);
  const float2 eh = mul2(e, iq2), fh = mul2(f, iq2), hh = mul2(h, iq2);
  const float2 nchh = neg2(mul2(ch, hh));
  const float2 A = ffma2(bh, eh, nchh);
  const float2 Bm = ffma2(ah, fh, nchh);
  const float2 Cc = ffma2(bh, hh, neg2(mul2(ch, fh)));
  const float2 Dd = ffma2(ah, hh, neg2(mul2(ch, eh)));
  const float2 T = add2(A, Bm);
  const float2 amb = sub2(A, Bm);
  const float2 disc_raw = ffma2(amb, amb, mul2(bcast2(4.0f), mul2(Cc, Dd)));
  const float2 disc = make_float2((disc_raw.x < eps) ? eps : disc_raw.x, (disc_raw.y < eps) ? eps : disc_raw.y);
  const float2 ir = rsq2(disc);
  const float2 r = mul2(disc, ir);
  const float2 hT = mul2(bcast2(0.5f), T);
  const float2 l1r = ffma2(bcast2(-0.5f), r, hT), l2r = ffma2(bcast2(0.5f), r, hT);
  const float2 l1 = make_float2((l1r.x < 1.0f) ? 1.0f : l1r.x, (l1r.y < 1.0f) ? 1.0f : l1r.y);
  const float2 l2 = make_float2((l2r.x < 1.0f) ? 1.0f : l2r.x, (l2r.y < 1.0f) ? 1.0f : l2r.y);
  const float2 L1 = mul2(bcast2(kLn2), lg22(l1)), L2 = mul2(bcast2(kLn2), lg22(l2));
  const float2 arg = ffma2(L1, L1, ffma2(L2, L2, eps2));
  const float2 inv_d = rsq2(arg);
  const float2 d = mul2(arg, inv_d);
  if (WANT_SR || WANT_HR) {
    const float2 q1 = mul2(mul2(L1, inv_d), rcp2(l1)), q2 = mul2(mul2(L2, inv_d), rcp2(l2));
    // clamp sub-gradients (pass where raw >= 1); x*0 keeps NaN like torch
    const float2 dl1 = make_float2((l1r.x >= 1.0f) ? q1.x : l1r.x * 0.0f, (l1r.y >= 1.0f) ? q1.y : l1r.y * 0.0f);
    const float2 dl2 = make_float2((l2r.x >= 1.0f) ? q2.x : l2r.x * 0.0f, (l2r.y >= 1.0f) ? q2.y : l2r.y * 0.0f);
    const float2 dr = mul2(bcast2(0.5f), sub2(dl2, dl1));
    const float2 dq = mul2(mul2(bcast2(0.5f), dr), ir);
    const float2 ddisc = make_float2((disc_raw.x >= eps) ? dq.x : disc_raw.x * 0.0f,
                                     (disc_raw.y >= eps) ? dq.y : disc_raw.y * 0.0f);
    const float2 dT = ffma2(mul2(bcast2(2.0f), T), ddisc, mul2(bcast2(0.5f), add2(dl1, dl2)));
    const float2 dd4 = mul2(bcast2(4.0f), ddisc);
    const float2 dA = ffma2(neg2(dd4), Bm, dT);
    const float2 dB = ffma2(neg2(dd4), A, dT);
    const float2 dC = mul2(dd4, Dd);
    const float2 dD = mul2(dd4, Cc);
    if (WANT_SR) {
      const float2 dah = ffma2(dB, fh, mul2(dD, hh));
      const float2 dbh = ffma2(dA, eh, mul2(dC, hh));
      const float2 dch = neg2(ffma2(add2(dA, dB), hh, ffma2(dC, fh, mul2(dD, eh))));
      if (normalize) {
        const float2 s = ffma2(a, dah, ffma2(b, dbh, mul2(c, dch)));
        const float2 ddet = mul2(mul2(mul2(bcast2(-0.5f), s), mul2(iq1, iq1)), iq1);
        G.da = ffma2(dah, iq1, mul2(ddet, b));
        G.db = ffma2(dbh, iq1, mul2(ddet, a));
        G.dc = ffma2(dch, iq1, mul2(mul2(bcast2(-2.0f), ddet), c));
      } else {
        G.da = dah; G.db = dbh; G.dc = dch;
      }
    }
    if (WANT_HR) {
      const float2 deh = ffma2(dA, bh, neg2(mul2(dD, ch)));
      const float2 dfh = ffma2(dB, ah, neg2(mul2(dC, ch)));
      const float2 dhh = ffma2(neg2(add2(dA, dB)), ch, ffma2(dC, bh, mul2(dD, ah)));
      if (normalize) {
        const float2 s = ffma2(e, deh, ffma2(f, dfh, mul2(h, dhh)));
        const float2 ddet = mul2(mul2(mul2(bcast2(-0.5f), s), mul2(iq2, iq2)), iq2);
        G.de = ffma2(deh, iq2, mul2(ddet, f));
        G.df = ffma2(dfh, iq2, mul2(ddet, e));
        G.dh = ffma2(dhh, iq2, mul2(mul2(bcast2(-2.0f), ddet), h));
      } else {
        G.de = deh; G.df = dfh; G.dh = dhh;
      }
    }
  }
  return d;
}

// Chain + stores for one thread's 2 x 4 pixels; returns the sum of the valid distances.
template <bool WANT_HR>
SRST_DEV float st_chain_store(const float2 (&S1)[3][4], const float2 (&S2)[3][4], bool norm, float eps,
                              float* __restrict__ ds_sr, float* __restrict__ ds_hr, size_t img_off, int H, int W,
                              int gy0, int gx0, bool vec4) {
  float lsum = 0.f;
  float2 gs[3][4], gh[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    StPixelGrad2 G;
    G.da = G.db = G.dc = G.de = G.df = G.dh = make_float2(0.f, 0.f);
    const float2 d = st_pixel2<true, WANT_HR>(S1[0][j], S1[1][j], S1[2][j], S2[0][j], S2[1][j], S2[2][j], norm, eps, G);
    const bool okx = gx0 + j < W;
    lsum += (okx && gy0 < H) ? d.x : 0.f;
    lsum += (okx && gy0 + 1 < H) ? d.y : 0.f;
    gs[0][j] = G.da; gs[1][j] = G.db; gs[2][j] = G.dc;
    if (WANT_HR) { gh[0][j] = G.de; gh[1][j] = G.df; gh[2][j] = G.dh; }
  }
  if (gx0 >= W) return lsum;
  const size_t plane = (size_t)H * W;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    if (gy0 + hf >= H) continue;
    const size_t o = img_off + (size_t)(gy0 + hf) * W + gx0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v0 = hf ? gs[c][0].y : gs[c][0].x, v1 = hf ? gs[c][1].y : gs[c][1].x;
      const float v2 = hf ? gs[c][2].y : gs[c][2].x, v3 = hf ? gs[c][3].y : gs[c][3].x;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
      if (WANT_HR) {
        w0 = hf ? gh[c][0].y : gh[c][0].x; w1 = hf ? gh[c][1].y : gh[c][1].x;
        w2 = hf ? gh[c][2].y : gh[c][2].x; w3 = hf ? gh[c][3].y : gh[c][3].x;
      }
      if (vec4) {
        if (ds_sr) st4(ds_sr + o + c * plane, make_float4(v0, v1, v2, v3));
        if (WANT_HR) st4(ds_hr + o + c * plane, make_float4(w0, w1, w2, w3));
      } else {
        const float vv[4] = {v0, v1, v2, v3}, ww[4] = {w0, w1, w2, w3};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (gx0 + j < W) {
            if (ds_sr) ds_sr[o + c * plane + j] = vv[j];
            if (WANT_HR) ds_hr[o + c * plane + j] = ww[j];
          }
        }
      }
    }
  }
  return lsum;
}

// ------------------------------------------------------------------------------------------------
// Shared building blocks (row-pair interleaved planes: float2 at rp*PITCH + 2*col)
// ------------------------------------------------------------------------------------------------

// Load the RGB pixels of rows [gy0, gy0+ROWS) x cols [gx0, gx0+COLS) of image `base` ([3,H,W]),
// convert to grayscale, store row-pair interleaved into sG; zero outside the image (the reference
// zero-pads: padding='same', utils.py:219-222).  ROWS even; gx0 and COLS multiples of 4.
// DEPTH items (6 x LDG.128 each) are in flight per thread before the first conversion.
// If `gray_out` ([H][W] plane of this image) is given, the pixels of the tile interior
// rows [iy0, iy1) x cols [ix0, ix1) are also written there (once per pixel across tiles).
template <int ROWS, int COLS, int PITCH, int NT, int DEPTH, bool PX = false>
SRST_DEV void load_gray_tile(float* sG, const float* __restrict__ base, int H, int W, int gy0, int gx0,
                             bool vec4, int tid, float* __restrict__ gray_out = nullptr, int iy0 = 0, int iy1 = 0,
                             int ix0 = 0, int ix1 = 0, const float* __restrict__ px_other = nullptr,
                             float* px_acc = nullptr) {
  constexpr int C4 = COLS / 4;
  constexpr int NITEM = (ROWS / 2) * C4;
  const size_t plane = (size_t)H * W;
  if (vec4) {
#pragma unroll 1
    for (int it0 = tid; it0 < NITEM; it0 += DEPTH * NT) {
      float4 R[DEPTH][2], Gc[DEPTH][2], Bc[DEPTH][2];
      bool ok[DEPTH][2];
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) {
        const int it = it0 + u * NT;
        const int q = it / C4, c4 = it - q * C4;
        const int gx = gx0 + 4 * c4;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int gy = gy0 + 2 * q + hf;
          ok[u][hf] = (it < NITEM) && gy >= 0 && gy < H && gx >= 0 && gx < W;  // W % 4 == 0: all-in or all-out
          if (ok[u][hf]) {
            const float* p = base + (size_t)gy * W + gx;
            R[u][hf] = ldg4(p); Gc[u][hf] = ldg4(p + plane); Bc[u][hf] = ldg4(p + 2 * plane);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) {
        const int it = it0 + u * NT;
        if (it >= NITEM) break;
        const int q = it / C4, c4 = it - q * C4;
        float v[2][4];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (ok[u][hf]) {
            v[hf][0] = gray_of(R[u][hf].x, Gc[u][hf].x, Bc[u][hf].x);
            v[hf][1] = gray_of(R[u][hf].y, Gc[u][hf].y, Bc[u][hf].y);
            v[hf][2] = gray_of(R[u][hf].z, Gc[u][hf].z, Bc[u][hf].z);
            v[hf][3] = gray_of(R[u][hf].w, Gc[u][hf].w, Bc[u][hf].w);
          } else {
            v[hf][0] = v[hf][1] = v[hf][2] = v[hf][3] = 0.f;
          }
        }
        float* o = sG + q * PITCH + 8 * c4;
        st4(o, make_float4(v[0][0], v[1][0], v[0][1], v[1][1]));
        st4(o + 4, make_float4(v[0][2], v[1][2], v[0][3], v[1][3]));
        if (gray_out) {
          const int gx = gx0 + 4 * c4;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int gy = gy0 + 2 * q + hf;
            if (ok[u][hf] && gy >= iy0 && gy < iy1 && gx >= ix0 && gx < ix1)
              st4(gray_out + (size_t)gy * W + gx, make_float4(v[hf][0], v[hf][1], v[hf][2], v[hf][3]));
          }
        }
        if (PX && px_other) {  // fused Pixel term: squared RGB difference to the other image, tile interior only
          const int gx = gx0 + 4 * c4;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int gy = gy0 + 2 * q + hf;
            if (ok[u][hf] && gy >= iy0 && gy < iy1 && gx >= ix0 && gx < ix1) {
              const float* po = px_other + (size_t)gy * W + gx;
              const float4 a = ldg4(po), b = ldg4(po + plane), c = ldg4(po + 2 * plane);
              float sacc = 0.f;
              float d;
              d = R[u][hf].x - a.x; sacc = fmaf(d, d, sacc); d = R[u][hf].y - a.y; sacc = fmaf(d, d, sacc);
              d = R[u][hf].z - a.z; sacc = fmaf(d, d, sacc); d = R[u][hf].w - a.w; sacc = fmaf(d, d, sacc);
              d = Gc[u][hf].x - b.x; sacc = fmaf(d, d, sacc); d = Gc[u][hf].y - b.y; sacc = fmaf(d, d, sacc);
              d = Gc[u][hf].z - b.z; sacc = fmaf(d, d, sacc); d = Gc[u][hf].w - b.w; sacc = fmaf(d, d, sacc);
              d = Bc[u][hf].x - c.x; sacc = fmaf(d, d, sacc); d = Bc[u][hf].y - c.y; sacc = fmaf(d, d, sacc);
              d = Bc[u][hf].z - c.z; sacc = fmaf(d, d, sacc); d = Bc[u][hf].w - c.w; sacc = fmaf(d, d, sacc);
              *px_acc += sacc;
            }
          }
        }
      }
    }
  } else {
    for (int it = tid; it < NITEM; it += NT) {
      const int q = it / C4, c4 = it - q * C4;
      const int gx = gx0 + 4 * c4;
      float v[2][4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int gy = gy0 + 2 * q + hf;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = gx + j;
          v[hf][j] = 0.f;
          if (gy >= 0 && gy < H && x >= 0 && x < W) {
            const float* p = base + (size_t)gy * W + x;
            v[hf][j] = gray_of(__ldg(p), __ldg(p + plane), __ldg(p + 2 * plane));
            if (gray_out && gy >= iy0 && gy < iy1 && x >= ix0 && x < ix1) gray_out[(size_t)gy * W + x] = v[hf][j];
            if (PX && px_other && gy >= iy0 && gy < iy1 && x >= ix0 && x < ix1) {
              const float* po = px_other + (size_t)gy * W + x;
              const float d0 = __ldg(p) - __ldg(po), d1 = __ldg(p + plane) - __ldg(po + plane);
              const float d2 = __ldg(p + 2 * plane) - __ldg(po + 2 * plane);
              *px_acc += fmaf(d0, d0, fmaf(d1, d1, d2 * d2));
            }
          }
        }
      }
      float* o = sG + q * PITCH + 8 * c4;
      st4(o, make_float4(v[0][0], v[1][0], v[0][1], v[1][1]));
      st4(o + 4, make_float4(v[0][2], v[1][2], v[0][3], v[1][3]));
    }
  }
}

// Asynchronous staging of the raw RGB rows [gy0, gy0+ROWS) x cols [gx0, gx0+COLS) of one image
// into shared memory ([3][ROWS][COLS] row-major) with 16-byte LDGSTS copies: no registers are
// held while the data is in flight, every copy of the tile is outstanding at once (one exposed
// round trip per tile), and pixels outside the image are zero-filled by the copy itself.
template <int ROWS, int COLS, int NT>
SRST_DEV void stage_rgb_async(float* sS, const float* __restrict__ base, int H, int W, int gy0, int gx0, int tid) {
  constexpr int C4 = COLS / 4;
  const size_t plane = (size_t)H * W;
  for (int it = tid; it < 3 * ROWS * C4; it += NT) {
    const int c4 = it % C4, rc = it / C4;
    const int r = rc % ROWS, c = rc / ROWS;
    const int gy = gy0 + r, gx = gx0 + 4 * c4;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;  // W % 4 == 0: all-in or all-out
    cp_async16(sS + (c * ROWS + r) * COLS + 4 * c4, ok ? base + c * plane + (size_t)gy * W + gx : base, ok);
  }
  cp_async_commit();
}

// Staged RGB -> grayscale, row-pair interleaved (loss.py:400-401).
template <int ROWS, int COLS, int PITCH, int NT>
SRST_DEV void convert_staged_gray(float* sG, const float* sS, int tid) {
  constexpr int C4 = COLS / 4;
  for (int it = tid; it < (ROWS / 2) * C4; it += NT) {
    const int q = it / C4, c4 = it - q * C4;
    float v[2][4];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float* p = sS + (2 * q + hf) * COLS + 4 * c4;
      const float4 R = ld4(p), Gc = ld4(p + ROWS * COLS), Bc = ld4(p + 2 * ROWS * COLS);
      v[hf][0] = gray_of(R.x, Gc.x, Bc.x);
      v[hf][1] = gray_of(R.y, Gc.y, Bc.y);
      v[hf][2] = gray_of(R.z, Gc.z, Bc.z);
      v[hf][3] = gray_of(R.w, Gc.w, Bc.w);
    }
    float* o = sG + q * PITCH + 8 * c4;
    st4(o, make_float4(v[0][0], v[1][0], v[0][1], v[1][1]));
    st4(o + 4, make_float4(v[0][2], v[1][2], v[0][3], v[1][3]));
  }
}

// L2 prefetch of the rows [gy0, gy0+ROWS) x cols [gx0, gx0+COLS) of the three planes of an image
// (one prefetch per 128-byte line): issued for the HR tile before the SR tile is processed.
template <int ROWS, int COLS, int NT>
SRST_DEV void prefetch_tile_l2(const float* __restrict__ base, int H, int W, int gy0, int gx0, int tid) {
#ifndef SRST_EMULATE
  constexpr int L = (COLS + 31) / 32 + 1;
  const size_t plane = (size_t)H * W;
  for (int it = tid; it < ROWS * 3 * L; it += NT) {
    const int l = it % L, rc = it / L;
    const int c = rc % 3, r = rc / 3;
    const int gy = gy0 + r, gx = gx0 + 32 * l;
    if (gy >= 0 && gy < H && gx < W && gx + 31 >= 0) {
      const float* p = base + c * plane + (size_t)gy * W + max(gx, 0);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
#endif
}

// Gaussian-derivative filter pair on a row pair, CS output columns:
//   ox = (A * dg|) * g-     oy = (B * g|) * dg-       (utils.py:219-222; A == B == gray forward)
// pa / pb point at (first input row pair, window start column) of planes A / B; the window is WIN
// columns wide and the CS outputs are centred at window columns CEN..CEN+CS-1.  RG+1 input row pairs
// feed the two output rows through the tap pairs dgp / gp.
template <int RG, int CS, int WIN, int CEN, bool SAME, class Taps>
SRST_DEV void grad_rowpair_rows(const float* const (&pa)[RG + 1], const float* const (&pb)[RG + 1], const Taps& tp,
                                float2 (&ox)[CS], float2 (&oy)[CS]) {
  static_assert(RG % 2 == 0 && WIN % 2 == 0, "row-pair gradient needs an even radius");
  float2 tA[WIN], tB[WIN];
#pragma unroll
  for (int j = 0; j < WIN; ++j) { tA[j] = make_float2(0.f, 0.f); tB[j] = make_float2(0.f, 0.f); }
#pragma unroll
  for (int q = 0; q <= RG; ++q) {
    float2 va[WIN], vb[WIN];
#pragma unroll
    for (int m = 0; m < WIN / 2; ++m) {
      const float4 t = ld4(pa[q] + 4 * m);
      va[2 * m] = make_float2(t.x, t.y);
      va[2 * m + 1] = make_float2(t.z, t.w);
      if (!SAME) {
        const float4 u = ld4(pb[q] + 4 * m);
        vb[2 * m] = make_float2(u.x, u.y);
        vb[2 * m + 1] = make_float2(u.z, u.w);
      }
    }
#pragma unroll
    for (int j = CEN - RG; j < CEN + CS + RG; ++j) {
      const float2 a = va[j], b = SAME ? va[j] : vb[j];
      // input rows 2q (.x) and 2q+1 (.y) of the window
      tA[j] = ffma2(bcast2(a.x), tp.dgp[2 * q], tA[j]);
      tA[j] = ffma2(bcast2(a.y), tp.dgp[2 * q + 1], tA[j]);
      tB[j] = ffma2(bcast2(b.x), tp.gp[2 * q], tB[j]);
      tB[j] = ffma2(bcast2(b.y), tp.gp[2 * q + 1], tB[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < CS; ++j) {
    float2 sx = make_float2(0.f, 0.f), sy = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t <= 2 * RG; ++t) {
      sx = ffma2(tA[CEN + j + t - RG], bcast2(tp.g[t]), sx);
      if (t != RG) sy = ffma2(tB[CEN + j + t - RG], bcast2(tp.dg[t]), sy);  // dg[RG] == 0
    }
    ox[j] = sx;
    oy[j] = sy;
  }
}

// Same filter pair from a ROW-MAJOR plane (row pitch PITCH floats): `p` points at (first input row,
// window start column); rows 2q and 2q+1 are loaded separately and used as broadcast scalars.
template <int RG, int CS, int WIN, int CEN, int PITCH, class Taps>
SRST_DEV void grad_rowpair_rm(const float* p, const Taps& tp, float2 (&ox)[CS], float2 (&oy)[CS]) {
  static_assert(RG % 2 == 0 && WIN % 4 == 0, "row-major gradient window must be a multiple of 4 columns");
  float2 tA[WIN], tB[WIN];
#pragma unroll
  for (int j = 0; j < WIN; ++j) { tA[j] = make_float2(0.f, 0.f); tB[j] = make_float2(0.f, 0.f); }
#pragma unroll
  for (int q = 0; q <= RG; ++q) {
    float r0[WIN], r1[WIN];
#pragma unroll
    for (int m = 0; m < WIN / 4; ++m) {
      const float4 a = ld4(p + (2 * q) * PITCH + 4 * m), b = ld4(p + (2 * q + 1) * PITCH + 4 * m);
      r0[4 * m] = a.x; r0[4 * m + 1] = a.y; r0[4 * m + 2] = a.z; r0[4 * m + 3] = a.w;
      r1[4 * m] = b.x; r1[4 * m + 1] = b.y; r1[4 * m + 2] = b.z; r1[4 * m + 3] = b.w;
    }
#pragma unroll
    for (int j = CEN - RG; j < CEN + CS + RG; ++j) {
      tA[j] = ffma2(bcast2(r0[j]), tp.dgp[2 * q], tA[j]);
      tA[j] = ffma2(bcast2(r1[j]), tp.dgp[2 * q + 1], tA[j]);
      tB[j] = ffma2(bcast2(r0[j]), tp.gp[2 * q], tB[j]);
      tB[j] = ffma2(bcast2(r1[j]), tp.gp[2 * q + 1], tB[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < CS; ++j) {
    float2 sx = make_float2(0.f, 0.f), sy = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t <= 2 * RG; ++t) {
      sx = ffma2(tA[CEN + j + t - RG], bcast2(tp.g[t]), sx);
      if (t != RG) sy = ffma2(tB[CEN + j + t - RG], bcast2(tp.dg[t]), sy);
    }
    ox[j] = sx;
    oy[j] = sy;
  }
}

// Same, for planes whose RG+1 input row pairs are PITCH floats apart.
template <int RG, int CS, int WIN, int CEN, int PITCH, bool SAME, class Taps>
SRST_DEV void grad_rowpair(const float* pa, const float* pb, const Taps& tp, float2 (&ox)[CS], float2 (&oy)[CS]) {
  const float* ra[RG + 1];
  const float* rb[RG + 1];
#pragma unroll
  for (int q = 0; q <= RG; ++q) { ra[q] = pa + q * PITCH; rb[q] = pb + q * PITCH; }
  grad_rowpair_rows<RG, CS, WIN, CEN, SAME>(ra, rb, tp, ox, oy);
}

// Horizontal rho-pass on a row pair, 4 output columns: out[j] = sum_t k[t] * v[CEN + j + t - RK].
template <int RK, int CS, int WIN, int CEN, class Taps>
SRST_DEV void smooth_h_rowpair_n(const float* row, const Taps& tp, float2 (&out)[CS]) {
  float2 v[WIN];
#pragma unroll
  for (int m = 0; m < WIN / 2; ++m) {
    const float4 t = ld4(row + 4 * m);
    v[2 * m] = make_float2(t.x, t.y);
    v[2 * m + 1] = make_float2(t.z, t.w);
  }
#pragma unroll
  for (int j = 0; j < CS; ++j) {
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t <= 2 * RK; ++t) s = ffma2(v[CEN + j + t - RK], bcast2(tp.k[t]), s);
    out[j] = s;
  }
}

template <int RK, int WIN, int CEN, class Taps>
SRST_DEV void smooth_h_rowpair(const float* row, const Taps& tp, float2 (&out)[4]) {
  float2 v[WIN];
#pragma unroll
  for (int m = 0; m < WIN / 2; ++m) {
    const float4 t = ld4(row + 4 * m);
    v[2 * m] = make_float2(t.x, t.y);
    v[2 * m + 1] = make_float2(t.z, t.w);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t <= 2 * RK; ++t) s = ffma2(v[CEN + j + t - RK], bcast2(tp.k[t]), s);
    out[j] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Forward
//
// Persistent, warp-specialised CTA.  NC compute threads run the filter phases; NP producer threads
// (one or two warps) only move data: they read the RGB tile (+halo) of the NEXT unit -- a unit is
// one image of one tile, SR then HR -- convert it to grayscale and park it in shared memory while
// the compute warps are still smoothing the current unit, so no compute warp ever waits on HBM/L2.
// Hand-off uses named barriers: FULL (producer arrives, consumers sync) and EMPTY (consumers
// arrive once the gradient phase has consumed the gray tile, producer syncs).
// ------------------------------------------------------------------------------------------------
constexpr int kBarFull = 1 /* +buffer (1,2) */, kBarCompute = 3, kBarEmpty = 4 /* +buffer (4,5) */;

template <int TH_, int TW_, int RS_, int CSB_, int NP_, int RG_, int RK_, int MINB_, bool STAGE_ = false>
struct StFwdCfg {
  static constexpr int TH = TH_, TW = TW_, RS = RS_, CSB = CSB_, NP = NP_, RG = RG_, RK = RK_, MINB = MINB_;
  static constexpr bool STAGE = STAGE_;  // stage raw RGB with cp.async over D|V instead of LDG -> registers
  static constexpr int LDEPTH = 2;       // gray-tile items (6 x LDG.128 each) in flight per thread
  static constexpr int NC = TH * TW / 8;     // one horizontal-pass item (2 rows x 4 cols) per compute thread
  static constexpr int NT = NC + NP;
  static constexpr int HXD = round_up4(RK);  // x halo of the gradient (D) and V regions
  static constexpr int OFF = round_up4(RG);
  static constexpr int HXG = HXD + OFF;      // x halo of the gray (G) region
  static constexpr int GH = TH + 2 * (RG + RK), GW = TW + 2 * HXG, PG = smem_pitch(2 * GW);
  static constexpr int DH = TH + 2 * RK, DW = TW + 2 * HXD, PD = smem_pitch(2 * DW);
  static constexpr int PV = PD;
  static constexpr int NSEG = TH / RS;
  static constexpr int BW_LO = (OFF - RG) / 2 * 2, BW_HI = (OFF + CSB + RG + 1) / 2 * 2, BWIN = BW_HI - BW_LO;
  static constexpr int DW_LO = (HXD - RK) / 2 * 2, DW_HI = (HXD + 4 + RK + 1) / 2 * 2, DWIN = DW_HI - DW_LO;
  static constexpr int D_FLOATS = (DH / 2) * PD, V_FLOATS = (TH / 2) * PV, G_FLOATS = (GH / 2) * PG;
  // smem: D (Ix, Iy) | V (3 planes) | G (gray).  With producer warps or RGB staging the gray tile
  // needs its own buffer; otherwise it aliases V (it is dead once the gradient phase is done).
  static constexpr bool G_ALIAS = (NP == 0) && !STAGE;
  static constexpr int G_OFF = 2 * D_FLOATS + (G_ALIAS ? 0 : 3 * V_FLOATS);
  static constexpr int NGBUF = NP > 0 ? 2 : 1;  // the producer runs a whole unit ahead: two gray buffers
  static constexpr int SMEM_FLOATS = G_ALIAS ? 2 * D_FLOATS + cmax(3 * V_FLOATS, G_FLOATS) : G_OFF + NGBUF * G_FLOATS;
  static constexpr size_t SMEM_BYTES = sizeof(float) * SMEM_FLOATS;
  static_assert(!STAGE || 3 * GH * GW <= 2 * D_FLOATS + 3 * V_FLOATS, "RGB staging does not fit over the D|V region");
  static_assert((CSB == 4 || CSB == 8) && DW % CSB == 0, "bad gradient segment width");
  static_assert(TH % RS == 0 && RS % 2 == 0 && TH % 2 == 0 && TW % 4 == 0 && RG % 2 == 0 && RK % 2 == 0,
                "bad forward tile");
  static_assert(NC % 32 == 0 && NP % 32 == 0 && NT <= 1024, "bad forward block size");
};

// Compute-warp phases B-D for one unit: consumes the gray tile the producer parked in sG and leaves
// the smoothed tensor (Jxx,Jyy,Jxy) of this thread's 8 pixels (rows 2q, 2q+1; cols ox0..ox0+3 of the
// tile) in S[3][4] (.x = even row, .y = odd row).
template <class C, bool PX, class Taps>
SRST_DEV void st_unit_tensor(float* smem, const float* __restrict__ base, float* __restrict__ gray_out, bool vec4, int H,
                             int W, int y0, int x0, const Taps& tp, int tid, int gbuf, bool release_gray,
                             float2 (&S)[3][4], const float* __restrict__ px_other = nullptr, float* px_acc = nullptr) {
  float* sD0 = smem;
  float* sD1 = sD0 + C::D_FLOATS;
  float* sV = sD1 + C::D_FLOATS;
  float* sG = smem + C::G_OFF + gbuf * C::G_FLOATS;

  if (C::NP > 0) {
    bar_sync(kBarFull + gbuf, C::NT);  // gray tile of this unit is in sG (also: every compute thread is done with sV)
  } else {
    bar_sync(kBarCompute, C::NC);  // every thread is done with sD / sV of the previous unit
    if (C::STAGE && vec4) {
      stage_rgb_async<C::GH, C::GW, C::NC>(smem, base, H, W, y0 - (C::RG + C::RK), x0 - C::HXG, tid);
      cp_async_wait_all();
      bar_sync(kBarCompute, C::NC);
      convert_staged_gray<C::GH, C::GW, C::PG, C::NC>(sG, smem, tid);
    } else {
      load_gray_tile<C::GH, C::GW, C::PG, C::NC, C::LDEPTH, PX>(sG, base, H, W, y0 - (C::RG + C::RK), x0 - C::HXG, vec4,
                                                                tid, gray_out, y0, y0 + C::TH, x0, x0 + C::TW, px_other,
                                                                px_acc);
    }
    bar_sync(kBarCompute, C::NC);
  }

  // Phase B: Ix, Iy on the D region; forced to zero outside the image because the reference
  // zero-pads the *products* for the rho-smoothing (utils.py:225-230).
  for (int it = tid; it < (C::DH / 2) * (C::DW / C::CSB); it += C::NC) {
    const int seg = it / (C::DH / 2), q = it - seg * (C::DH / 2);
    const int dx0 = C::CSB * seg;
    const int gy = y0 - C::RK + 2 * q, gx0 = x0 - C::HXD + dx0;
    float2 Ix[C::CSB], Iy[C::CSB];
    if (gy + 1 >= 0 && gy < H && gx0 + C::CSB - 1 >= 0 && gx0 < W) {
      const float* p = sG + q * C::PG + 2 * (dx0 + C::BW_LO);
      grad_rowpair<C::RG, C::CSB, C::BWIN, C::OFF - C::BW_LO, C::PG, true>(p, p, tp, Ix, Iy);
      const bool r0 = gy >= 0, r1 = gy + 1 < H;  // gy < H and gy + 1 >= 0 hold here
#pragma unroll
      for (int j = 0; j < C::CSB; ++j) {
        const bool ok = (gx0 + j >= 0) && (gx0 + j < W);
        Ix[j].x = (ok && r0) ? Ix[j].x : 0.f;
        Ix[j].y = (ok && r1) ? Ix[j].y : 0.f;
        Iy[j].x = (ok && r0) ? Iy[j].x : 0.f;
        Iy[j].y = (ok && r1) ? Iy[j].y : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < C::CSB; ++j) { Ix[j] = make_float2(0.f, 0.f); Iy[j] = make_float2(0.f, 0.f); }
    }
    float* o0 = sD0 + q * C::PD + 2 * dx0;
    float* o1 = sD1 + q * C::PD + 2 * dx0;
#pragma unroll
    for (int j = 0; j < C::CSB; j += 2) {
      st4(o0 + 2 * j, make_float4(Ix[j].x, Ix[j].y, Ix[j + 1].x, Ix[j + 1].y));
      st4(o1 + 2 * j, make_float4(Iy[j].x, Iy[j].y, Iy[j + 1].x, Iy[j + 1].y));
    }
  }
  if (C::NP > 0 && release_gray) bar_arrive(kBarEmpty + gbuf, C::NT);  // this gray buffer may be refilled (unit + 2)
  bar_sync(kBarCompute, C::NC);

  // Phase C: vertical rho-pass of the three products; a lane owns one column and RS output rows
  // (RS/2 row pairs).  Columns outside the image hold zeros and are only cleared.
  {
    const int c_lo = max(0, C::HXD - x0), c_hi = min(C::DW, W + C::HXD - x0);  // valid D columns
    const int ncol = c_hi - c_lo;
    const int nclr = C::DW - ncol;
    for (int it = tid; it < nclr * (C::TH / 2); it += C::NC) {
      const int q = it / nclr, u = it - q * nclr;
      const int dx = (u < c_lo) ? u : (c_hi + (u - c_lo));
#pragma unroll
      for (int c = 0; c < 3; ++c) st2(sV + c * C::V_FLOATS + q * C::PV + 2 * dx, make_float2(0.f, 0.f));
    }
    for (int it = tid; it < ncol * C::NSEG; it += C::NC) {
      const int seg = it / ncol, dx = c_lo + (it - seg * ncol);
      float2 acc[3][C::RS / 2];
#pragma unroll
      for (int j = 0; j < C::RS / 2; ++j) {
        acc[0][j] = make_float2(0.f, 0.f); acc[1][j] = make_float2(0.f, 0.f); acc[2][j] = make_float2(0.f, 0.f);
      }
      const float* p0 = sD0 + (seg * (C::RS / 2)) * C::PD + 2 * dx;
      const float* p1 = sD1 + (seg * (C::RS / 2)) * C::PD + 2 * dx;
#pragma unroll
      for (int rq = 0; rq < C::RS / 2 + C::RK; ++rq) {
        const float2 ix = ld2(p0 + rq * C::PD), iy = ld2(p1 + rq * C::PD);
        const float2 pxx = mul2(ix, ix), pyy = mul2(iy, iy), pxy = mul2(ix, iy);  // packed FMUL2
#pragma unroll
        for (int jp = 0; jp < C::RS / 2; ++jp) {
          const int u0 = 2 * rq - 2 * jp;  // tap-pair index of input row 2rq for output pair jp
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
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float* o = sV + c * C::V_FLOATS + (seg * (C::RS / 2)) * C::PV + 2 * dx;
#pragma unroll
        for (int jp = 0; jp < C::RS / 2; ++jp) st2(o + jp * C::PV, acc[c][jp]);
      }
    }
  }
  bar_sync(kBarCompute, C::NC);

  // Phase D: horizontal rho-pass; a lane owns 4 consecutive columns of one row pair.
  {
    const int seg = tid / (C::TH / 2), q = tid - seg * (C::TH / 2);
    const int ox0 = 4 * seg;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      smooth_h_rowpair<C::RK, C::DWIN, C::HXD - C::DW_LO>(sV + c * C::V_FLOATS + q * C::PV + 2 * (ox0 + C::DW_LO), tp,
                                                          S[c]);
  }
  // no barrier here: the next unit starts with bar_sync(kBarFull) over all compute threads, which
  // orders these sV reads before the next phase C and the sD reads of phase C before the next phase B
}

template <class C, bool PX = false>
__global__ void __launch_bounds__(C::NT, C::MINB)
st_forward_kernel(const __grid_constant__ StFwdParams<C::RG, C::RK> P) {
  SRST_DYN_SMEM(float, smem);
  __shared__ float s_red[32];
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x;
  const int ntiles = P.B * P.tiles_y * P.tiles_x;
  pdl_wait();     // previous kernel of the stream is complete (it may have produced sr or used the workspace)
  pdl_trigger();  // the next PDL-launched kernel may start launching; it waits for this grid itself

  if (C::NP > 0 && tid >= C::NC) {
    // ---- producer warps: RGB tile (+halo) -> gray -> sG, one unit ahead of the compute warps ----
    const int ptid = tid - C::NC;
    int u = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int t = tile;
      const int tx = t % P.tiles_x;
      t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int b = t / P.tiles_y;
      const size_t img_off = (size_t)b * 3 * P.H * P.W;
#pragma unroll 1
      for (int img = 0; img < 2; ++img, ++u) {
        const int gb = u & 1;
        if (u >= 2) bar_sync(kBarEmpty + gb, C::NT);  // compute warps consumed the tile of unit u-2
        load_gray_tile<C::GH, C::GW, C::PG, (C::NP > 0 ? C::NP : 32), 3>(
            smem + C::G_OFF + gb * C::G_FLOATS, (img ? P.hr : P.sr) + img_off, P.H, P.W,
            ty * C::TH - (C::RG + C::RK), tx * C::TW - C::HXG, P.vec4 != 0, ptid,
            (img ? P.gray_hr : P.gray_sr) ? (img ? P.gray_hr : P.gray_sr) + (size_t)b * P.H * P.W : nullptr, ty * C::TH,
            ty * C::TH + C::TH, tx * C::TW, tx * C::TW + C::TW);
        __threadfence_block();
        bar_arrive(kBarFull + gb, C::NT);
      }
    }
    return;
  }

  // ---- compute warps ----
  const bool want_hr = P.ds_hr != nullptr;
  const bool norm = P.normalize != 0;
  const int seg = tid / (C::TH / 2), q = tid - seg * (C::TH / 2);
  float lsum = 0.f;
  [[maybe_unused]] float pxsum = 0.f;
  // NP > 0: persistent loop over tiles (the producer runs one unit ahead); NP == 0: one tile per CTA
  int tile = blockIdx.x;
  do {
    int t = tile;
    const int tx = t % P.tiles_x;
    t /= P.tiles_x;
    const int ty = t % P.tiles_y;
    const int b = t / P.tiles_y;
    const int y0 = ty * C::TH, x0 = tx * C::TW;
    const size_t img_off = (size_t)b * 3 * P.H * P.W;
    const bool last_tile = tile + (int)gridDim.x >= ntiles;  // units of the last tile have no unit + 2

    float2 S1[3][4], S2[3][4];
    const size_t gray_off = (size_t)b * P.H * P.W;
    st_unit_tensor<C, false>(smem, P.sr + img_off, P.gray_sr ? P.gray_sr + gray_off : nullptr, P.vec4 != 0, P.H, P.W, y0,
                             x0, P.taps, tid, 0, !last_tile, S1);
    if constexpr (PX)  // fused Pixel term: the HR unit's loader also reads the SR pixels of the tile interior
      st_unit_tensor<C, true>(smem, P.hr + img_off, P.gray_hr ? P.gray_hr + gray_off : nullptr, P.vec4 != 0, P.H, P.W, y0,
                              x0, P.taps, tid, (C::NP > 0 ? 1 : 0), !last_tile, S2, P.sr + img_off, &pxsum);
    else
      st_unit_tensor<C, false>(smem, P.hr + img_off, P.gray_hr ? P.gray_hr + gray_off : nullptr, P.vec4 != 0, P.H, P.W, y0,
                               x0, P.taps, tid, (C::NP > 0 ? 1 : 0), !last_tile, S2);

    // Per-pixel chain on this thread's 2 x 4 pixels, then the ds stores.
    if (want_hr)
      lsum += st_chain_store<true>(S1, S2, norm, P.eps, P.ds_sr, P.ds_hr, img_off, P.H, P.W, y0 + 2 * q, x0 + 4 * seg,
                                   P.vec4 != 0);
    else
      lsum += st_chain_store<false>(S1, S2, norm, P.eps, P.ds_sr, P.ds_hr, img_off, P.H, P.W, y0 + 2 * q, x0 + 4 * seg,
                                    P.vec4 != 0);
  } while (C::NP > 0 && (tile += (int)gridDim.x) < ntiles);

  // Deterministic loss reduction: block partial -> workspace; the last block to finish sums all
  // partials in a fixed order (double) and re-zeroes the workspace for the next call.
  [[maybe_unused]] __shared__ float s_red_px[PX ? 32 : 1];
  lsum = warp_sum(lsum);
  if constexpr (PX) pxsum = warp_sum(pxsum);
  if ((tid & 31) == 0) {
    s_red[tid >> 5] = lsum;
    if constexpr (PX) s_red_px[tid >> 5] = pxsum;
  }
  bar_sync(kBarCompute, C::NC);
  if (tid == 0) {
    float bs = 0.f;
    for (int w = 0; w < C::NC / 32; ++w) bs += s_red[w];
    P.partials[blockIdx.x] = bs;
    if constexpr (PX) {
      float bp = 0.f;
      for (int w = 0; w < C::NC / 32; ++w) bp += s_red_px[w];
      P.px_partials[blockIdx.x] = bp;
    }
    __threadfence();
    const unsigned int t = atomicAdd(P.ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  bar_sync(kBarCompute, C::NC);
  if (s_last) {
    __threadfence();
    if (tid < 32) {
      double acc = 0.0;
      [[maybe_unused]] double accp = 0.0;
      for (unsigned int i = tid; i < gridDim.x; i += 32) {
        acc += (double)__ldcg(P.partials + i);
        P.partials[i] = 0.f;
        if constexpr (PX) {
          accp += (double)__ldcg(P.px_partials + i);
          P.px_partials[i] = 0.f;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if constexpr (PX) accp += __shfl_xor_sync(0xffffffffu, accp, o);
      }
      if (tid == 0) {
        P.loss_out[0] = (float)(acc * (double)P.inv_count);
        if constexpr (PX) P.loss_out[1] = (float)(accp * (double)P.inv_count / 3.0);  // mean over B*3*H*W
        *P.ticket = 0u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------------
template <int TH_, int TW_, int RS_, int NT_, int RG_, int RK_, int MINB_, int CSD_ = 4>
struct StBwdCfg {
  static constexpr int TH = TH_, TW = TW_, RS = RS_, NT = NT_, RG = RG_, RK = RK_, MINB = MINB_;
  static constexpr int CSD = CSD_;  // output columns per phase-D' item: 8 reads a 24-column window per 8 outputs
                                    // (3x shared-memory amplification) instead of 20 per 4 (5x)
  static constexpr int HXE = round_up4(RG);  // x halo of the E region (where dIx, dIy are needed)
  static constexpr int HXK = round_up4(RK);
  static constexpr int EH = TH + 2 * RG, EW = TW + 2 * HXE, PE = smem_pitch(2 * EW);
  static constexpr int GH = EH + 2 * RG, GW = EW + 2 * HXE, PG = smem_pitch(2 * GW);  // gray region
  static constexpr int VW = EW + 2 * HXK, PV = smem_pitch(2 * VW);                     // vertical-pass output
  static constexpr int SH = EH + 2 * RK;                                              // staged ds rows
  static constexpr int NSEG = EH / RS;
  static constexpr int BW_LO = (HXE - RG) / 2 * 2, BW_HI = (HXE + 4 + RG + 1) / 2 * 2, BWIN = BW_HI - BW_LO;
  static constexpr int DW_LO = (HXK - RK) / 2 * 2, DW_HI = (HXK + CSD + RK + 1) / 2 * 2, DWIN = DW_HI - DW_LO;
  // smem: V (3 planes) | I (Ix, Iy) | G (gray) | X: staged ds planes [3][SH][VW] row-major, later
  // dIx|dIy (the staging is dead once the vertical pass has consumed it)
  static constexpr int V_FLOATS = (EH / 2) * PV, I_FLOATS = (EH / 2) * PE, G_FLOATS = (GH / 2) * PG;
  static constexpr int S_FLOATS = SH * VW;
  static constexpr int X_FLOATS = cmax(2 * I_FLOATS, 3 * S_FLOATS);
  static constexpr int G_OFF = (3 * V_FLOATS + 2 * I_FLOATS + 31) / 32 * 32;             // 128-byte aligned (TMA destination)
  static constexpr int X_OFF = (G_OFF + G_FLOATS + 31) / 32 * 32;
  static constexpr int RMW_LO = (HXE - RG) / 4 * 4, RMWIN = (HXE + 4 + RG + 3) / 4 * 4 - RMW_LO;  // row-major gradient window
  static_assert(GH * GW <= G_FLOATS, "row-major gray tile must fit the gray region");
  static constexpr int SMEM_FLOATS = X_OFF + X_FLOATS;
  static constexpr size_t SMEM_BYTES = sizeof(float) * SMEM_FLOATS;
  static_assert(EH % RS == 0 && RS % 2 == 0 && TH % 2 == 0 && TW % 4 == 0 && RG % 2 == 0 && RK % 2 == 0,
                "bad backward tile");
  static_assert(NT % 32 == 0 && NT <= 1024, "bad backward block size");
  static_assert((CSD == 4 || CSD == 8) && EW % CSD == 0, "bad phase-D' item width");
};

template <class C, bool PX = false>
__global__ void __launch_bounds__(C::NT, C::MINB)
st_backward_kernel(const __grid_constant__ StBwdParams<C::RG, C::RK> P) {
  SRST_DYN_SMEM(float, smem);
  float* sV = smem;                    // [3][EH/2][PV]  vertical rho-pass of ds
  float* sI0 = sV + 3 * C::V_FLOATS;   // Ix [EH/2][PE]
  float* sI1 = sI0 + C::I_FLOATS;      // Iy
  float* sG = smem + C::G_OFF;         // gray [GH/2][PG] (row pairs), or [GH][GW] row-major when it comes by TMA
  float* sX = smem + C::X_OFF;         // staged ds [3][SH][VW], later dIx|dIy
  float* sS = sX;
  float* sdI0 = sX;
  float* sdI1 = sX + C::I_FLOATS;

  const int tid = threadIdx.x;
  // First-wave CTAs can become resident while the previous kernel of the stream (the forward, in a
  // training step's backward pass) is still draining.  Their gray tile and Ix, Iy depend on `img` only,
  // which that kernel does not write, so they are built first; ds (written by the forward) and
  // grad_out (written by the autograd op right before this launch) are touched after the wait.
  const bool early = P.use_tma && !P.use_gray && (int)blockIdx.x < P.early_ctas;
  if (!early) pdl_wait();
  pdl_trigger();
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
  const int xv0 = x0 - C::HXE - C::HXK;        // first staged / V column
  const int yv0 = y0 - C::RG - C::RK;          // first staged row

  // Stage the three ds planes (+halo).  Preferred path: ONE TMA box copy issued by one thread
  // (rows/columns outside the image are zero-filled by the hardware = the zero padding of the
  // adjoint smoothing); otherwise 16-byte cp.async copies, or scalar loads for unaligned tensors.
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ __align__(8) unsigned long long s_mbar_g;
  // the gray box is needed first (phase B'): queue it ahead of the much larger ds box
  if (P.use_gray && tid == 0)
    tma_stage_begin(&s_mbar_g, sG, &P.gray_map, x0 - 2 * C::HXE, y0 - 2 * C::RG, b, C::GW, C::GH, 1);
  if (P.use_tma) {
    if (!early && tid == 0) tma_stage_begin(&s_mbar, sS, &P.ds_map, xv0, yv0, b * 3, C::VW, C::SH, 3);
  } else {
    const float* dsb = P.ds + img_off;
    constexpr int C4 = C::VW / 4;
    if (P.vec4) {
      // (plane c, row r, 16-byte column group c4) walked incrementally: no div/mod in the loop
      constexpr int DC4 = C::NT % C4, DR = C::NT / C4;
      static_assert(DR + 1 < C::SH, "staging walk assumes at most one row wrap per step");
      int c4 = tid % C4, r = tid / C4, c = 0;
      while (r >= C::SH) { r -= C::SH; ++c; }
      while (c < 3) {
        const int gy = yv0 + r, gx = xv0 + 4 * c4;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        cp_async16(sS + (c * C::SH + r) * C::VW + 4 * c4, ok ? dsb + c * plane + (size_t)gy * W + gx : dsb, ok);
        c4 += DC4; r += DR;
        if (c4 >= C4) { c4 -= C4; ++r; }
        if (r >= C::SH) { r -= C::SH; ++c; }
      }
      cp_async_commit();
    } else {
      for (int it = tid; it < 3 * C::SH * C::VW; it += C::NT) {
        const int vx = it % C::VW, rc = it / C::VW;
        const int r = rc % C::SH, c = rc / C::SH;
        const int gy = yv0 + r, gx = xv0 + vx;
        sS[it] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(dsb + c * plane + (size_t)gy * W + gx) : 0.f;
      }
    }
  }

  // Phase A': gray tile of the image (halo 2*RG rows, 2*HXE cols): one TMA box of the gray plane
  // the forward saved, or RGB loads + conversion.
  if (P.use_gray) {
    __syncthreads();  // the barrier is initialised before anyone polls it
    tma_stage_wait(&s_mbar_g);
  } else {
    load_gray_tile<C::GH, C::GW, C::PG, C::NT, 1>(sG, P.img + img_off, H, W, y0 - 2 * C::RG, x0 - 2 * C::HXE,
                                                  P.vec4 != 0, tid);
    __syncthreads();
  }

  // Phase B': recompute Ix, Iy on the E region (zero outside the image).
  for (int it = tid; it < (C::EH / 2) * (C::EW / 4); it += C::NT) {
    const int seg = it / (C::EH / 2), q = it - seg * (C::EH / 2);
    const int ex0 = 4 * seg;
    const int gy = y0 - C::RG + 2 * q, gx0 = x0 - C::HXE + ex0;
    float2 Ix[4], Iy[4];
    if (gy + 1 >= 0 && gy < H && gx0 + 3 >= 0 && gx0 < W) {
      if (P.use_gray) {
        grad_rowpair_rm<C::RG, 4, C::RMWIN, C::HXE - C::RMW_LO, C::GW>(sG + (2 * q) * C::GW + ex0 + C::RMW_LO, tp, Ix, Iy);
      } else {
        const float* p = sG + q * C::PG + 2 * (ex0 + C::BW_LO);
        grad_rowpair<C::RG, 4, C::BWIN, C::HXE - C::BW_LO, C::PG, true>(p, p, tp, Ix, Iy);
      }
      const bool r0 = gy >= 0, r1 = gy + 1 < H;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (gx0 + j >= 0) && (gx0 + j < W);
        Ix[j].x = (ok && r0) ? Ix[j].x : 0.f;
        Ix[j].y = (ok && r1) ? Ix[j].y : 0.f;
        Iy[j].x = (ok && r0) ? Iy[j].x : 0.f;
        Iy[j].y = (ok && r1) ? Iy[j].y : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { Ix[j] = make_float2(0.f, 0.f); Iy[j] = make_float2(0.f, 0.f); }
    }
    float* o0 = sI0 + q * C::PE + 2 * ex0;
    float* o1 = sI1 + q * C::PE + 2 * ex0;
    st4(o0, make_float4(Ix[0].x, Ix[0].y, Ix[1].x, Ix[1].y));
    st4(o0 + 4, make_float4(Ix[2].x, Ix[2].y, Ix[3].x, Ix[3].y));
    st4(o1, make_float4(Iy[0].x, Iy[0].y, Iy[1].x, Iy[1].y));
    st4(o1 + 4, make_float4(Iy[2].x, Iy[2].y, Iy[3].x, Iy[3].y));
  }
  if (early) {
    pdl_wait();  // the previous kernel of the stream is complete: ds may be fetched now
    if (tid == 0) tma_stage_begin(&s_mbar, sS, &P.ds_map, xv0, yv0, b * 3, C::VW, C::SH, 3);
    __syncthreads();  // the barrier is initialised before anyone polls it
  }
  if (P.use_tma) tma_stage_wait(&s_mbar); else cp_async_wait_all();
  __syncthreads();  // Ix, Iy complete; staged ds planes have landed

  // Phase C': vertical rho-pass of the three staged ds planes; a lane owns one column and RS rows.
  // The adjoint of the zero-padded symmetric smoothing is the same zero-padded smoothing.
  {
    const int c_lo = max(0, -xv0), c_hi = min(C::VW, W - xv0);
    const int ncol = c_hi - c_lo;
    const int nclr = C::VW - ncol;
    for (int it = tid; it < nclr * (C::EH / 2); it += C::NT) {
      const int q = it / nclr, u = it - q * nclr;
      const int vx = (u < c_lo) ? u : (c_hi + (u - c_lo));
#pragma unroll
      for (int c = 0; c < 3; ++c) st2(sV + c * C::V_FLOATS + q * C::PV + 2 * vx, make_float2(0.f, 0.f));
    }
    // one item = (plane, row segment, column): the three ds planes are independent, so splitting them
    // over threads costs nothing and keeps every lane busy (ncol * NSEG alone is < NT for most tiles)
    for (int it = tid; it < 3 * ncol * C::NSEG; it += C::NT) {
      const int cs = it / ncol, vx = c_lo + (it - cs * ncol);
      const int c = cs / C::NSEG, seg = cs - c * C::NSEG;
      float2 acc[C::RS / 2];
#pragma unroll
      for (int j = 0; j < C::RS / 2; ++j) acc[j] = make_float2(0.f, 0.f);
      const float* p = sS + (c * C::SH + seg * C::RS) * C::VW + vx;
#pragma unroll
      for (int r = 0; r < C::RS + 2 * C::RK; ++r) {
        const float v = p[r * C::VW];
#pragma unroll
        for (int jp = 0; jp < C::RS / 2; ++jp) {
          const int u = r - 2 * jp;
          if (u >= 0 && u <= 2 * C::RK + 1) acc[jp] = ffma2(bcast2(v), tp.kp[u], acc[jp]);
        }
      }
      float* o = sV + c * C::V_FLOATS + (seg * (C::RS / 2)) * C::PV + 2 * vx;
#pragma unroll
      for (int jp = 0; jp < C::RS / 2; ++jp) st2(o + jp * C::PV, acc[jp]);
    }
  }
  __syncthreads();  // the staging is dead from here on: sdI may overwrite it

  // Phase D': horizontal rho-pass -> E = K*ds at the E-region pixels, then the product rule
  //   dIx = 2 Ix Exx + Iy Exy ,  dIy = 2 Iy Eyy + Ix Exy      (adjoint of utils.py:225-229)
  for (int it = tid; it < (C::EH / 2) * (C::EW / C::CSD); it += C::NT) {
    const int seg = it / (C::EH / 2), q = it - seg * (C::EH / 2);
    const int ex0 = C::CSD * seg;
    float2 E[3][C::CSD];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      smooth_h_rowpair_n<C::RK, C::CSD, C::DWIN, C::HXK - C::DW_LO>(
          sV + c * C::V_FLOATS + q * C::PV + 2 * (ex0 + C::DW_LO), tp, E[c]);
    float* o0 = sdI0 + q * C::PE + 2 * ex0;
    float* o1 = sdI1 + q * C::PE + 2 * ex0;
#pragma unroll
    for (int m = 0; m < C::CSD / 2; ++m) {  // two columns per LDS.128 / STS.128
      const float4 ixv = ld4(sI0 + q * C::PE + 2 * ex0 + 4 * m), iyv = ld4(sI1 + q * C::PE + 2 * ex0 + 4 * m);
      const float2 ix[2] = {make_float2(ixv.x, ixv.y), make_float2(ixv.z, ixv.w)};
      const float2 iy[2] = {make_float2(iyv.x, iyv.y), make_float2(iyv.z, iyv.w)};
      float2 dx[2], dy[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * m + h;
        const float2 ix2 = make_float2(2.0f * ix[h].x, 2.0f * ix[h].y), iy2 = make_float2(2.0f * iy[h].x, 2.0f * iy[h].y);
        dx[h] = ffma2(ix2, E[0][j], make_float2(iy[h].x * E[2][j].x, iy[h].y * E[2][j].y));
        dy[h] = ffma2(iy2, E[1][j], make_float2(ix[h].x * E[2][j].x, ix[h].y * E[2][j].y));
      }
      st4(o0 + 4 * m, make_float4(dx[0].x, dx[0].y, dx[1].x, dx[1].y));
      st4(o1 + 4 * m, make_float4(dy[0].x, dy[0].y, dy[1].x, dy[1].y));
    }
  }
  __syncthreads();

  // Phase E': adjoint of the derivative filters.  The adjoint of a zero-padded correlation is the
  // correlation with flipped taps; g is symmetric and dg antisymmetric, so
  //   dgray = -[ (dIx * dg|) * g-  +  (dIy * g|) * dg- ]
  // i.e. the forward gradient operators applied to dIx and dIy, negated.
  const float scale = -__ldg(P.grad_out) * P.inv_count;
  [[maybe_unused]] const float px_scale = PX ? __ldg(P.grad_px) * P.inv_count * (2.0f / 3.0f) : 0.f;
  for (int it = tid; it < (C::TH / 2) * (C::TW / 4); it += C::NT) {
    const int seg = it / (C::TH / 2), q = it - seg * (C::TH / 2);
    const int ox0 = 4 * seg;
    const int gy = y0 + 2 * q, gx0 = x0 + ox0;
    if (gy >= H || gx0 >= W) continue;
    float2 ax[4], ay[4];
    grad_rowpair<C::RG, 4, C::BWIN, C::HXE - C::BW_LO, C::PE, false>(sdI0 + q * C::PE + 2 * (ox0 + C::BW_LO),
                                                                   sdI1 + q * C::PE + 2 * (ox0 + C::BW_LO), tp, ax, ay);
    const float coef[3] = {kGrayR, kGrayG, kGrayB};
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      if (gy + hf >= H) continue;
      float dgr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) dgr[j] = ((hf ? ax[j].y : ax[j].x) + (hf ? ay[j].y : ay[j].x)) * scale;
      float* o = P.d_img + img_off + (size_t)(gy + hf) * W + gx0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (P.vec4) {
          float4 r = make_float4(coef[c] * dgr[0], coef[c] * dgr[1], coef[c] * dgr[2], coef[c] * dgr[3]);
          if constexpr (PX) {  // fused Pixel term: d MSE / d img = 2 (img - other) / (3 B H W)
            const size_t e = img_off + (size_t)(gy + hf) * W + gx0 + c * plane;
            const float4 a = ldg4(P.img + e), b = ldg4(P.px_other + e);
            r.x = fmaf(px_scale, a.x - b.x, r.x); r.y = fmaf(px_scale, a.y - b.y, r.y);
            r.z = fmaf(px_scale, a.z - b.z, r.z); r.w = fmaf(px_scale, a.w - b.w, r.w);
          }
          st4(o + c * plane, r);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gx0 + j < W) {
              float r = coef[c] * dgr[j];
              if constexpr (PX) {
                const size_t e = img_off + (size_t)(gy + hf) * W + gx0 + j + c * plane;
                r = fmaf(px_scale, __ldg(P.img + e) - __ldg(P.px_other + e), r);
              }
              o[c * plane + j] = r;
            }
        }
      }
    }
  }
}

}  // namespace srst
