/* CPU oracle for the Best-Buddy loss with an ARBITRARY patch geometry -- TEST INFRASTRUCTURE ONLY
 * (see oracle/bbg_oracle.py).
 *
 * Plain-C restatement of SebastianBitsch/SRGAN-ST loss.py:115-141 for any (ksize, pad, stride), fp32 with one fixed
 * operation order (the order of the CUDA path in srgan_st_b200/csrc/bb_generic.cuh, so indices compare bit-exactly):
 *   patches  : F.unfold(kernel_size=k, padding=p, stride=s): element c*k*k + ky*k + kx of patch py*nx + px is
 *              img[c][py*s - p + ky][px*s - p + kx], zero outside the image; ny = (H + 2p - k)/s + 1   (loss.py:116-129)
 *   distance : l2 = max((|x|^2 + |y|^2) - 2 x.y, 0) (utils.py:173-187), l1 = sum |x - y| (utils.py:166-172)
 *   score    : alpha*d(sr_i, y_j) + beta*d(gt_i, y_j), argmin = first minimal index              (loss.py:132-135)
 *   loss     : mean |sr_patch - y[argmin]| or mean square                                       (loss.py:139)
 *   d_sr     : the criterion's gradient folded back over the (possibly overlapping) patches
 * Norms and dot products are accumulated with fmaf over the patch elements in ascending order.
 * The pyramid levels come from the caller (oracle/bb_oracle.c bb_oracle_pyramid or the reference's own).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int npatch(int size, int k, int p, int s) {
  const int span = size + 2 * p - k;
  return span < 0 ? 0 : span / s + 1;
}

static void read_patch(const float* img, int H, int W, int k, int p, int s, int nx, int n, float* v) {
  const int py = n / nx, px = n % nx;
  for (int c = 0; c < 3; ++c)
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const int y = py * s - p + ky, x = px * s - p + kx;
        v[(c * k + ky) * k + kx] = (y >= 0 && y < H && x >= 0 && x < W) ? img[((size_t)c * H + y) * W + x] : 0.f;
      }
}
static float normd(const float* v, int d) {
  float n = 0.f;
  for (int i = 0; i < d; ++i) n = fmaf(v[i], v[i], n);
  return n;
}
static float dotd(const float* a, const float* b, int d) {
  float r = 0.f;
  for (int i = 0; i < d; ++i) r = fmaf(a[i], b[i], r);
  return r;
}

/* criterion: 0 = L1, 1 = L2 (mean square); | 0x100 = dist_norm 'l1'.  d_sr may be NULL. */
int bbg_oracle_forward(const float* sr, const float* gt, const float* gt2, const float* gt4, int B, int H, int W, int k,
                       int p, int s, float alpha, float beta, int criterion, int64_t* idx, double* loss_out,
                       float* best_out, float* second_out, float* d_sr) {
  const int dist_l1 = (criterion & 0x100) != 0;
  criterion &= 0xff;
  const int D = 3 * k * k;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  const int n0x = npatch(W, k, p, s), N0 = npatch(H, k, p, s) * n0x;
  const int n2x = npatch(W2, k, p, s), N2 = npatch(H2, k, p, s) * n2x;
  const int n4x = npatch(W4, k, p, s), N4 = npatch(H4, k, p, s) * n4x;
  if (N0 <= 0 || N2 <= 0 || N4 <= 0) return 2; /* F.unfold raises on an empty level */
  const int N = N0, M = N0 + N2 + N4;
  float* q1 = malloc(sizeof(float) * (size_t)N * D);
  float* q2 = malloc(sizeof(float) * (size_t)N * D);
  float* y = malloc(sizeof(float) * (size_t)M * D);
  float* xn = malloc(sizeof(float) * N);
  float* gn = malloc(sizeof(float) * N);
  float* yn = malloc(sizeof(float) * M);
  if (!q1 || !q2 || !y || !xn || !gn || !yn) return 1;
  double total = 0.0;
  const float scale = 1.0f / ((float)B * (float)N * (float)D);
  for (int b = 0; b < B; ++b) {
    const float* s0 = sr + (size_t)b * 3 * H * W;
    const float* g0 = gt + (size_t)b * 3 * H * W;
    const float* g2 = gt2 + (size_t)b * 3 * H2 * W2;
    const float* g4 = gt4 + (size_t)b * 3 * H4 * W4;
    for (int i = 0; i < N; ++i) {
      read_patch(s0, H, W, k, p, s, n0x, i, q1 + (size_t)i * D);
      xn[i] = normd(q1 + (size_t)i * D, D);
      read_patch(g0, H, W, k, p, s, n0x, i, q2 + (size_t)i * D);
      gn[i] = normd(q2 + (size_t)i * D, D);
    }
    for (int j = 0; j < M; ++j) {
      if (j < N0) read_patch(g0, H, W, k, p, s, n0x, j, y + (size_t)j * D);
      else if (j < N0 + N2) read_patch(g2, H2, W2, k, p, s, n2x, j - N0, y + (size_t)j * D);
      else read_patch(g4, H4, W4, k, p, s, n4x, j - N0 - N2, y + (size_t)j * D);
      yn[j] = normd(y + (size_t)j * D, D);
    }
    if (d_sr)
      for (size_t e = 0; e < (size_t)3 * H * W; ++e) d_sr[(size_t)b * 3 * H * W + e] = 0.f;
    for (int i = 0; i < N; ++i) {
      const float *a1 = q1 + (size_t)i * D, *a2 = q2 + (size_t)i * D;
      float best = INFINITY, second = INFINITY;
      int bi = 0;
      for (int j = 0; j < M; ++j) {
        const float* c = y + (size_t)j * D;
        float d1, d2;
        if (dist_l1) {
          d1 = 0.f; d2 = 0.f;
          for (int e = 0; e < D; ++e) { d1 = d1 + fabsf(a1[e] - c[e]); d2 = d2 + fabsf(a2[e] - c[e]); }
        } else {
          d1 = fmaf(-2.0f, dotd(a1, c, D), xn[i] + yn[j]);
          d2 = fmaf(-2.0f, dotd(a2, c, D), gn[i] + yn[j]);
          d1 = d1 < 0.f ? 0.f : d1; /* torch.clamp(min=0) keeps NaN */
          d2 = d2 < 0.f ? 0.f : d2;
        }
        const float sa = alpha * d1, sb = beta * d2;
        const float sc = sa + sb;
        if (sc < best || (sc != sc && best == best)) { second = best; best = sc; bi = j; }
        else if (sc < second) second = sc;
      }
      idx[(size_t)b * N + i] = bi;
      if (best_out) best_out[(size_t)b * N + i] = best;
      if (second_out) second_out[(size_t)b * N + i] = second;
      const int py = i / n0x, px = i % n0x;
      for (int e = 0; e < D; ++e) {
        const float d = a1[e] - y[(size_t)bi * D + e];
        total += criterion == 0 ? fabs((double)d) : (double)d * (double)d;
        if (d_sr) {
          const int c = e / (k * k), ky = (e / k) % k, kx = e % k;
          const int yy = py * s - p + ky, xx = px * s - p + kx;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const float gv = criterion == 0 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * d;
            d_sr[(((size_t)b * 3 + c) * H + yy) * W + xx] += gv * scale;
          }
        }
      }
    }
  }
  *loss_out = total / ((double)B * N * D);
  free(q1); free(q2); free(y); free(xn); free(gn); free(yn);
  return 0;
}
