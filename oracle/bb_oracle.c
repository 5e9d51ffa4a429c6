/* CPU oracle for the Best-Buddy loss -- TEST INFRASTRUCTURE ONLY (see oracle/bb_oracle.py).
 *
 * Plain-C restatement of the reference's algorithm (SebastianBitsch/SRGAN-ST loss.py:115-141,
 * utils.py:173-187), in fp32 with one fixed operation order, so that argmin indices can be
 * compared bit-exactly with the CUDA path:
 *   patches   : F.unfold(k=3, stride=3, pad=0): element c*9+ky*3+kx, patch py*(W/3)+px  (loss.py:116-121)
 *   pyramid   : bicubic (A=-0.75), align_corners=False, scale 1/2 and 1/4 of gt itself  (loss.py:123,127)
 *   distance  : (|x|^2 + |y|^2) - 2 x.y, clamped at 0                                   (utils.py:183-187)
 *   score     : alpha*d(sr_i, y_j) + beta*d(gt_i, y_j)                                  (loss.py:132-133)
 *   argmin    : first minimal index (torch.min)                                         (loss.py:135)
 *   loss      : mean |sr_patch - y[argmin]| (L1) or mean square (L2)                    (loss.py:139)
 * The 27-term dot products and norms are accumulated with fmaf in k = 0..26 order; torch's bmm uses
 * an unspecified order, so against the reference itself indices are compared under the near-tie
 * protocol of tests/test_bb_*.py.  Build: make -C oracle  (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define D 27

static float cubic4(float a, float b, float c, float d) {
  const float w0 = -0.09375f, w1 = 0.59375f; /* cubic convolution weights at t = 0.5, A = -0.75 */
  return ((a * w0 + b * w1) + c * w1) + d * w0;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* gt [planes,H,W] -> o2 [planes,H/2,W/2], o4 [planes,H/4,W/4] */
void bb_oracle_pyramid(const float* gt, int planes, int H, int W, float* o2, float* o4) {
  for (int lvl = 0; lvl < 2; ++lvl) {
    const int s = lvl ? 4 : 2, Ho = H / s, Wo = W / s;
    float* o = lvl ? o4 : o2;
    for (int p = 0; p < planes; ++p)
      for (int y = 0; y < Ho; ++y)
        for (int x = 0; x < Wo; ++x) {
          const int sy = lvl ? 4 * y : 2 * y - 1, sx = lvl ? 4 * x : 2 * x - 1;
          float r[4];
          for (int i = 0; i < 4; ++i) {
            const float* row = gt + ((size_t)p * H + clampi(sy + i, 0, H - 1)) * W;
            r[i] = cubic4(row[clampi(sx, 0, W - 1)], row[clampi(sx + 1, 0, W - 1)], row[clampi(sx + 2, 0, W - 1)],
                          row[clampi(sx + 3, 0, W - 1)]);
          }
          o[((size_t)p * Ho + y) * Wo + x] = cubic4(r[0], r[1], r[2], r[3]);
        }
  }
}

static void read_patch(const float* img, int H, int W, int nx, int p, float* v) {
  const int py = p / nx, px = p % nx;
  for (int c = 0; c < 3; ++c)
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) v[c * 9 + ky * 3 + kx] = img[((size_t)c * H + 3 * py + ky) * W + 3 * px + kx];
}
static float normd(const float* v, int d) {
  float n = 0.f;
  for (int k = 0; k < d; ++k) n = fmaf(v[k], v[k], n);
  return n;
}
static float dotd(const float* a, const float* b, int d) {
  float s = 0.f;
  for (int k = 0; k < d; ++k) s = fmaf(a[k], b[k], s);
  return s;
}
/* GramLoss descriptor (reference loss.py:180-184 gram_matrix): patch viewed as features [3][9],
 * G = F F^T / 27, flattened row-major.  Same fixed order as the CUDA kernel. */
static void gram9(const float* v, float* g) {
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      float s = 0.f;
      for (int t = 0; t < 9; ++t) s = fmaf(v[a * 9 + t], v[b * 9 + t], s);
      g[a * 3 + b] = s / 27.0f;
    }
}
/* PatchwiseStructureTensorLoss descriptor (reference loss.py:325-345): the patch as a 3x3 image ->
 * Grayscale (loss.py:327) -> utils.structure_tensor with zero 'same' padding (utils.py:212-233; on a
 * 3x3 image only the five central taps of each filter matter) -> utils.normalize (utils.py:236-239).
 * taps = g5[5] | dg5[5] | k5[5] (offsets -2..2).  Explicitly rounded fp32, fixed order, same as the
 * CUDA kernel (bb_kernels.cuh pst_state). */
static const float* g_pst_taps = 0;
void bb_oracle_set_pst_taps(const float* taps15) { g_pst_taps = taps15; }
static void vert3(const float* w, const float* X, float* o) {
  for (int i = 0; i < 3; ++i)
    for (int x = 0; x < 3; ++x) {
      float s = 0.f;
      for (int j = 0; j < 3; ++j) s = fmaf(w[j - i + 2], X[j * 3 + x], s);
      o[i * 3 + x] = s;
    }
}
static void horz3(const float* w, const float* X, float* o) {
  for (int i = 0; i < 3; ++i)
    for (int x = 0; x < 3; ++x) {
      float s = 0.f;
      for (int j = 0; j < 3; ++j) s = fmaf(w[j - x + 2], X[i * 3 + j], s);
      o[i * 3 + x] = s;
    }
}
static void pst27(const float* v, float* out) {
  const float *g5 = g_pst_taps, *dg5 = g_pst_taps + 5, *k5 = g_pst_taps + 10;
  float gray[9], t[9], Ix[9], Iy[9], p[9], J[3][9];
  for (int i = 0; i < 9; ++i) {
    const float a = 0.2989f * v[i], b = 0.587f * v[9 + i], c = 0.114f * v[18 + i];
    const float ab = a + b;
    gray[i] = ab + c;
  }
  vert3(dg5, gray, t); horz3(g5, t, Ix);
  vert3(g5, gray, t);  horz3(dg5, t, Iy);
  for (int c = 0; c < 3; ++c) {
    for (int i = 0; i < 9; ++i) p[i] = (c == 1 ? Iy[i] : Ix[i]) * (c == 0 ? Ix[i] : Iy[i]);
    vert3(k5, p, t);
    horz3(k5, t, J[c]);
  }
  for (int i = 0; i < 9; ++i) {
    const float ab = J[0][i] * J[1][i], cc = J[2][i] * J[2][i];
    const float det = ab - cc;
    const float q = sqrtf(det + 1e-12f);
    for (int c = 0; c < 3; ++c) out[c * 9 + i] = J[c][i] / q;
  }
}
/* descriptor of a patch: mode 0 = raw 27 values (BestBuddyLoss), mode 1 = Gram matrix (GramLoss),
 * mode 2 = normalised structure tensor of the patch (PatchwiseStructureTensorLoss) */
static void describe(const float* img, int H, int W, int nx, int p, int mode, float* out) {
  float v[D];
  read_patch(img, H, W, nx, p, v);
  if (mode == 0) for (int k = 0; k < D; ++k) out[k] = v[k];
  else if (mode == 1) gram9(v, out);
  else pst27(v, out);
}
/* descriptors of all level-0 patches of img [B,3,H,W] -> out [B,N,Dd] (tests) */
void bb_oracle_describe(const float* img, int B, int H, int W, int mode, float* out) {
  const int Dd = mode == 1 ? 9 : D, nx = W / 3, N = (H / 3) * nx;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i) describe(img + (size_t)b * 3 * H * W, H, W, nx, i, mode, out + ((size_t)b * N + i) * Dd);
}

/* Returns 0 on success.  idx [B,N] int64; loss_out 1 double; best/second [B,N] fp32 scores (may be NULL). */
int bb_oracle_forward_mode(const float* sr, const float* gt, const float* gt2, const float* gt4, int B, int H, int W,
                           float alpha, float beta, int criterion, int mode, int64_t* idx, double* loss_out,
                           float* best_out, float* second_out) {
  const int Dd = mode == 1 ? 9 : D;
  const int dist_l1 = (criterion & 0x100) != 0; /* dist_norm='l1' (utils.py:166-172), flag as in include/srst.h */
  criterion &= 0xff;
  const int n0x = W / 3, N0 = (H / 3) * n0x;
  const int H2 = H / 2, W2 = W / 2, n2x = W2 / 3, N2 = (H2 / 3) * n2x;
  const int H4 = H / 4, W4 = W / 4, n4x = W4 / 3, N4 = (H4 / 3) * n4x;
  const int N = N0, M = N0 + N2 + N4;
  float* q1 = malloc(sizeof(float) * (size_t)N * Dd);
  float* q2 = malloc(sizeof(float) * (size_t)N * Dd);
  float* y = malloc(sizeof(float) * (size_t)M * Dd);
  float* xn = malloc(sizeof(float) * N);
  float* gn = malloc(sizeof(float) * N);
  float* yn = malloc(sizeof(float) * M);
  if (!q1 || !q2 || !y || !xn || !gn || !yn) return 1;
  double total = 0.0;
  for (int b = 0; b < B; ++b) {
    const float* s0 = sr + (size_t)b * 3 * H * W;
    const float* g0 = gt + (size_t)b * 3 * H * W;
    const float* g2 = gt2 + (size_t)b * 3 * H2 * W2;
    const float* g4 = gt4 + (size_t)b * 3 * H4 * W4;
    for (int i = 0; i < N; ++i) {
      describe(s0, H, W, n0x, i, mode, q1 + (size_t)i * Dd);
      xn[i] = normd(q1 + (size_t)i * Dd, Dd);
      describe(g0, H, W, n0x, i, mode, q2 + (size_t)i * Dd);
      gn[i] = normd(q2 + (size_t)i * Dd, Dd);
    }
    for (int j = 0; j < M; ++j) {
      if (j < N0) describe(g0, H, W, n0x, j, mode, y + (size_t)j * Dd);
      else if (j < N0 + N2) describe(g2, H2, W2, n2x, j - N0, mode, y + (size_t)j * Dd);
      else describe(g4, H4, W4, n4x, j - N0 - N2, mode, y + (size_t)j * Dd);
      yn[j] = normd(y + (size_t)j * Dd, Dd);
    }
    for (int i = 0; i < N; ++i) {
      float best = INFINITY, second = INFINITY;
      int bi = 0;
      for (int j = 0; j < M; ++j) {
        float d1 = fmaf(-2.0f, dotd(q1 + (size_t)i * Dd, y + (size_t)j * Dd, Dd), xn[i] + yn[j]);
        float d2 = fmaf(-2.0f, dotd(q2 + (size_t)i * Dd, y + (size_t)j * Dd, Dd), gn[i] + yn[j]);
        d1 = d1 < 0.f ? 0.f : d1; /* torch.clamp(min=0) keeps NaN (utils.py:187) */
        d2 = d2 < 0.f ? 0.f : d2;
        if (dist_l1) { /* sum_k |x_k - y_k| accumulated in k order, fp32 (utils.py:172) */
          d1 = 0.f; d2 = 0.f;
          for (int k = 0; k < Dd; ++k) {
            d1 = d1 + fabsf(q1[(size_t)i * Dd + k] - y[(size_t)j * Dd + k]);
            d2 = d2 + fabsf(q2[(size_t)i * Dd + k] - y[(size_t)j * Dd + k]);
          }
        }
        const float a = alpha * d1, bb = beta * d2;
        const float s = a + bb;
        /* torch.min (loss.py:135): first minimum; a NaN beats every number and the first NaN wins */
        if (s < best || (s != s && best == best)) { second = best; best = s; bi = j; }
        else if (s < second) second = s;
      }
      idx[(size_t)b * N + i] = bi;
      if (best_out) best_out[(size_t)b * N + i] = best;
      if (second_out) second_out[(size_t)b * N + i] = second;
      for (int k = 0; k < Dd; ++k) {
        const float d = q1[(size_t)i * Dd + k] - y[(size_t)bi * Dd + k];
        total += criterion == 0 ? fabs((double)d) : (double)d * (double)d;
      }
    }
  }
  *loss_out = total / ((double)B * N * Dd);
  free(q1); free(q2); free(y); free(xn); free(gn); free(yn);
  return 0;
}

int bb_oracle_forward(const float* sr, const float* gt, const float* gt2, const float* gt4, int B, int H, int W,
                      float alpha, float beta, int criterion, int64_t* idx, double* loss_out, float* best_out,
                      float* second_out) {
  return bb_oracle_forward_mode(sr, gt, gt2, gt4, B, H, W, alpha, beta, criterion, 0, idx, loss_out, best_out,
                                second_out);
}
