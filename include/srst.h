/* srst.h -- C ABI of libsrst.so: the B200-native structure-tensor / Best-Buddy loss hot path.
 *
 * This is the drop-in boundary for the loss hot path of SebastianBitsch/SRGAN-ST.  The reference
 * has no native code: the hot path is a chain of ATen calls behind two nn.Modules
 * (loss.py:380-413 StructureTensorLoss, loss.py:78-141 BestBuddyLoss).  Each entry point below
 * names the reference code it replaces.  The Python host (srgan_st_b200/loss.py) binds these with
 * ctypes from inside a torch.autograd.Function; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns int: 0 = ok, >0 = a cudaError_t value, <0 = an SRST_E_* code.
 *   - image/gradient buffers are DEVICE pointers to contiguous fp32 NCHW tensors owned by the
 *     caller; filter taps are HOST pointers (they are copied into kernel parameters).
 *   - no function allocates, synchronises or keeps state: all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*) and every scratch buffer comes in through `workspace`.
 *     The library is re-entrant across streams and devices as long as workspaces differ.
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef SRST_H_
#define SRST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRST_VERSION 204 /* major*100 + minor: the major number changes whenever an argument list changes */

#define SRST_E_INVALID (-1)     /* null pointer / non-positive size / bad enum */
#define SRST_E_UNSUPPORTED (-2) /* filter radius or patch geometry not compiled in */
#define SRST_E_WORKSPACE (-3)   /* workspace missing, misaligned or too small */
#define SRST_E_SHAPE (-4)       /* image shape not usable by this entry point */

int srst_version(void);
const char* srst_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * Structure-tensor loss.   Replaces, fused into one kernel per direction:
 *   torchvision Grayscale                        (loss.py:400-401)
 *   utils.structure_tensor  (10 conv2d / image)  (utils.py:212-233)
 *   utils.normalize / compute_invS1xS2           (utils.py:236-254)
 *   utils.compute_eigenvalues / compute_distance (utils.py:257-279)
 *   .mean() over pixels and batch                (loss.py:409,413)
 * and the autograd backward of that chain.
 *
 * g, dg: 2*r_sigma+1 host floats (utils.get_gaussian_kernel(sigma, also_dg=True), utils.py:194-208)
 * k    : 2*r_rho+1   host floats (utils.get_gaussian_kernel(rho))
 * Radii are padded with zero taps to the compiled classes r_sigma in {2,4}, r_rho in {4,8,12} (shared-memory
 * kernels; the reference default sigma=0.5, rho=2.0 is (2, 8)).  Larger radii, up to 64 each -- the reference's
 * radius max(int(4 sigma + 0.5), 1) is unbounded, utils.py:198 -- run on the generic-radius path: the same passes,
 * one thread per pixel, intermediates in global scratch planes (srst_st_workspace_bytes_r,
 * srst_st_backward_workspace_bytes, srst_st_backward_ws).
 * ------------------------------------------------------------------------------------------- */

/* 1 if the (r_sigma, r_rho) pair fits a compiled radius class, 2 if it takes the generic-radius path
 * (srst_st_forward + srst_st_backward_ws only; not the fused Pixel variant, not srst_st_features), else 0. */
int srst_st_supported(int r_sigma, int r_rho);

/* Tile-shape override for tests and tuning sweeps: force compiled forward / backward tile configuration
 * fwd_cfg / bwd_cfg (0 .. srst_st_num_cfgs(0|1) - 1) for the default radius class, -1 = the library's own
 * choice.  Process-wide; the environment variables SRST_ST_FWD_CFG / SRST_ST_BWD_CFG set the initial value
 * (read once). */
int srst_st_num_cfgs(int backward);
int srst_st_force_cfg(int fwd_cfg, int bwd_cfg);
/* Forward cfgs 0-2 are tiled shapes (32x64, 24x96 with two or three CTAs per SM), 3 and 4 the row-marching kernel
 * (96- / 112-column strips); backward cfgs: 0 = 28x56 tiles, 1 = 12x96 strips (both persistent).  The marching kernel cuts a strip into row chunks so
 * that small batches fill the machine; `blocks` > 0 forces the chunk height to blocks*16 rows (tests), <= 0 hands
 * the choice back to the library.  Process-wide; SRST_ST_CHUNK_BLOCKS sets the initial value. */
int srst_st_force_chunk_blocks(int blocks);

/* Scratch bytes srst_st_forward / srst_stpx_forward need for a [B,3,H,W] problem (per-CTA partial sums + ticket).
 * The workspace must be 16-byte aligned; its first 16 bytes (the ticket counter) must be ZERO before the
 * first use.  Every launch writes its partial sums before reading them and hands the ticket back zeroed,
 * so the buffer can be reused by later calls on the same stream; do not share one buffer between
 * streams, or between this entry point and the patch-loss entry points (they lay it out differently). */
size_t srst_st_workspace_bytes(int B, int H, int W);
/* Same, for a given radius pair: equal to srst_st_workspace_bytes() for the compiled classes; the generic-radius
 * path adds eleven [B,H,W] fp32 scratch planes (smoothed tensors of both images, one pass buffer, Ix/Iy when they
 * are not saved).  0 if the pair is not supported at all. */
size_t srst_st_workspace_bytes_r(int B, int H, int W, int r_sigma, int r_rho);
/* Scratch bytes srst_st_backward_ws needs: 0 for the compiled classes, five [B,H,W] fp32 planes on the
 * generic-radius path (no zero-initialisation required). */
size_t srst_st_backward_workspace_bytes(int B, int H, int W, int r_sigma, int r_rho);

/* Number of floats of a saved-gradient buffer ("ixy") for a [B,3,H,W] problem: [B][2][ceil(H/2)][W][2],
 * plane 0 = Ix (derivative along H), plane 1 = Iy, rows 2p / 2p+1 of a column interleaved (the layout the
 * kernels use in shared memory, so the backward fetches its tile with one TMA box copy). */
size_t srst_st_ixy_floats(int B, int H, int W);

/* Forward.  sr, hr: device [B,3,H,W] fp32.  Writes
 *   loss_out[0]  = mean over B*H*W pixels of the Riemannian distance   (device, 1 float)
 *   ds_sr        = d(sum of distances)/d(Jxx,Jyy,Jxy of SR), device [B,3,H,W], or NULL to skip
 *   ds_hr        = same w.r.t. the HR tensor, or NULL (only needed when hr requires grad)
 *   ixy_sr/_hr   = the Gaussian-derivative gradients Ix, Iy of the grayscale image (utils.py:219-222),
 *                  srst_st_ixy_floats() floats each, or NULL; required next to the matching ds_*
 * ds_* and ixy_* are the "saved intermediates" the backward pass consumes (20 B/pixel); ds is unscaled
 * (neither 1/(B*H*W) nor the upstream gradient is applied yet). */
int srst_st_forward(const float* sr, const float* hr, int B, int H, int W,
                    const float* g, const float* dg, int r_sigma,
                    const float* k, int r_rho,
                    int normalize, float eps,
                    float* loss_out, float* ds_sr, float* ds_hr, float* ixy_sr, float* ixy_hr,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Backward for one image tensor (sr or hr): ixy, ds = the matching ixy_* / ds_* buffers the forward
 * wrote; grad_out: device pointer to the upstream scalar gradient.  The image itself is not read again.
 * Writes d_img[B,3,H,W] = grad_out/(B*H*W) * dLoss_sum/dimg (adjoint smoothing, product rule,
 * adjoint Gaussian-derivative filters, grayscale weights).
 * Ordering: plain stream order -- every input may be produced by the kernel enqueued immediately before
 * this call on `stream` (the kernel waits for its predecessor before its first global read). */
int srst_st_backward(const float* ixy, const float* ds, const float* grad_out,
                     int B, int H, int W,
                     const float* g, const float* dg, int r_sigma,
                     const float* k, int r_rho,
                     float* d_img, void* stream);
/* Same with a scratch buffer: the only backward entry point of the generic-radius path (srst_st_backward returns
 * SRST_E_WORKSPACE there); for the compiled classes `workspace` is ignored and may be NULL.  On the generic path
 * the saved ixy buffer is planar [B][2][H][W] (written by srst_st_forward with the same radii). */
int srst_st_backward_ws(const float* ixy, const float* ds, const float* grad_out,
                        int B, int H, int W,
                        const float* g, const float* dg, int r_sigma,
                        const float* k, int r_rho,
                        float* d_img, void* workspace, size_t workspace_bytes, void* stream);

/* Structure-tensor FEATURES of one image tensor (diagnostic output; BASELINE.json north_star: "closed-form 2x2
 * eigendecomposition giving orientation, coherence and eigenvalues").  The reference computes these only in an
 * exploration notebook through the third-party `structure_tensor` package (data-exploration/structure_tensor.ipynb,
 * cells 12-18: eig_special_2d, arctan2 of the eigenvector, 1 - val[0]/val[1]), which is not part of its checkout:
 * the definitions below are this library's own (parity unpinned).
 *   J_out      [B,3,H,W]  Jxx, Jyy, Jxy: the smoothed tensor of utils.structure_tensor (utils.py:212-233)
 *   eig_out    [B,2,H,W]  lambda_small, lambda_large = (Jxx+Jyy)/2 -/+ sqrt(((Jxx-Jyy)/2)^2 + Jxy^2)
 *   orient_out [B,H,W]    1/2 atan2(2 Jxy, Jxx - Jyy): angle of the dominant-gradient eigenvector against the H axis
 *   coher_out  [B,H,W]    1 - lambda_small / lambda_large (0 where lambda_large == 0)
 * Any output may be NULL (at least one must be given).  No gradient. */
int srst_st_features(const float* img, int B, int H, int W,
                     const float* g, const float* dg, int r_sigma,
                     const float* k, int r_rho,
                     float* J_out, float* eig_out, float* orient_out, float* coher_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Structure-tensor loss with the "Pixel" MSE criterion fused in.  The warm-up
 * and training loops evaluate MSELoss(sr, gt) and StructureTensorLoss(sr, gt) on the same two
 * tensors one after the other (warmup.py:88-96, train.py:131-144, criteria registered in
 * config.py:71-93); here the squared RGB difference is accumulated while the forward kernel has
 * the HR tile in registers, and the backward kernel adds 2 (sr - hr) / (3 B H W) * grad_px to d_sr
 * in its final store -- no extra pass over either tensor.
 *   loss2_out[0] = structure-tensor loss (as srst_st_forward), loss2_out[1] = mean((sr - hr)^2)
 *   grad_st, grad_px: device scalars, the upstream gradients of the two terms.
 * Workspace: srst_st_workspace_bytes.
 * ------------------------------------------------------------------------------------------- */
int srst_stpx_forward(const float* sr, const float* hr, int B, int H, int W,
                      const float* g, const float* dg, int r_sigma,
                      const float* k, int r_rho,
                      int normalize, float eps,
                      float* loss2_out, float* ds_sr, float* ixy_sr,
                      void* workspace, size_t workspace_bytes, void* stream);
int srst_stpx_backward(const float* sr, const float* hr, const float* ixy, const float* ds,
                       const float* grad_st, const float* grad_px,
                       int B, int H, int W,
                       const float* g, const float* dg, int r_sigma,
                       const float* k, int r_rho,
                       float* d_sr, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Best-Buddy loss.   Replaces
 *   F.unfold x4, torch.cat                        (loss.py:116-130)
 *   utils.batch_pairwise_distance(...,'l2') x2    (utils.py:173-187; two torch.bmm)
 *   torch.min(score, dim=2), torch.gather         (loss.py:135-137)
 *   L1Loss / MSELoss                              (loss.py:139)
 * for ksize=3, stride=3, pad=0 (27-dim patches).  gt2 / gt4 are the bicubic x1/2 and x1/4
 * levels of gt ([B,3,H/2,W/2], [B,3,H/4,W/4]); pass NULL to have the library compute them with
 * the same taps F.interpolate(mode='bicubic', align_corners=False) uses (loss.py:123,127),
 * in which case `workspace` must also hold them (see srst_bb_workspace_bytes).
 * H, W >= 12; sizes that are not multiples of 3 / 12 follow the floor semantics of F.unfold and
 * F.interpolate(scale_factor=...).
 * ------------------------------------------------------------------------------------------- */
#define SRST_BB_L1 0
#define SRST_BB_L2 1
/* OR-ed into `criterion` of the three *_forward entry points below: search with dist_norm='l1'
 * (utils.py:166-172: alpha * sum|x - y| + beta * sum|g - y|) instead of the default squared l2 distance.  Every pair is
 * scored exactly then (no filter); the backward entry points ignore the flag (they only need the indices). */
#define SRST_BB_DIST_L1 0x100

size_t srst_bb_workspace_bytes(int B, int H, int W);

/* Forward: idx_out[B, N] int64 (N = (H/3)*(W/3)) = argmin over the M = N + N/4 + N/16 candidate
 * patches of alpha*d(sr_i, cand_j) + beta*d(gt_i, cand_j), first minimal index on ties
 * (torch.min rule); loss_out[0] = mean |sr_patch - cand[idx]| (L1) or mean square (L2). */
int srst_bb_forward(const float* sr, const float* gt, const float* gt2, const float* gt4,
                    int B, int H, int W, float alpha, float beta, int criterion,
                    int64_t* idx_out, float* loss_out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Backward: only the final criterion is differentiable w.r.t. sr (loss.py:139; argmin is not).
 * d_sr[B,3,H,W] = grad_out/(B*N*27) * sign(sr_patch - cand[idx])           (L1)
 *               = grad_out/(B*N*27) * 2*(sr_patch - cand[idx])             (L2)     */
int srst_bb_backward(const float* sr, const float* gt, const float* gt2, const float* gt4,
                     const int64_t* idx, const float* grad_out,
                     int B, int H, int W, int criterion,
                     float* d_sr, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gram loss (reference loss.py:146-225): the same three-level best-buddy search, but every 3x3x3
 * patch is replaced by its 3x3 Gram matrix (features [3][9], G = F F^T / 27; 9-dim descriptor) for
 * both the search and the final L1/MSE criterion.  Workspace: srst_bb_workspace_bytes.
 *   d_sr = through the criterion and the Gram descriptor of the SR patch (argmin is not differentiated).
 * ------------------------------------------------------------------------------------------- */
int srst_gram_forward(const float* sr, const float* gt, const float* gt2, const float* gt4,
                      int B, int H, int W, float alpha, float beta, int criterion,
                      int64_t* idx_out, float* loss_out,
                      void* workspace, size_t workspace_bytes, void* stream);
int srst_gram_backward(const float* sr, const float* gt, const float* gt2, const float* gt4,
                       const int64_t* idx, const float* grad_out,
                       int B, int H, int W, int criterion,
                       float* d_sr, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Patchwise structure-tensor loss (reference loss.py:292-375): the same three-level best-buddy
 * search, but every 3x3x3 patch is replaced by the det-normalised structure tensor of the patch
 * seen as a 3x3 image (Grayscale -> utils.structure_tensor(sigma, rho) with zero 'same' padding
 * -> utils.normalize; 27-dim descriptor c*9 + y*3 + x, c in Jxx,Jyy,Jxy), replacing
 *   vmap(vmap(s_norm)) over [B,N,3,3,3]            (loss.py:325-345)
 *   batch_pairwise_distance x2, torch.min, gather  (loss.py:361-369)
 *   L1Loss / MSELoss on the descriptors            (loss.py:371)
 * g, dg, k: host taps as for srst_st_forward; any radius >= 1 is accepted (a 3x3 image only sees
 * the five central taps).  Workspace: srst_bb_workspace_bytes.
 *   d_sr = through the criterion, utils.normalize and the structure tensor of the SR patch
 *          (argmin is not differentiated).
 * ------------------------------------------------------------------------------------------- */
int srst_pst_forward(const float* sr, const float* gt, const float* gt2, const float* gt4,
                     int B, int H, int W,
                     const float* g, const float* dg, int r_sigma, const float* k, int r_rho,
                     float alpha, float beta, int criterion,
                     int64_t* idx_out, float* loss_out,
                     void* workspace, size_t workspace_bytes, void* stream);
int srst_pst_backward(const float* sr, const float* gt, const float* gt2, const float* gt4,
                      const int64_t* idx, const float* grad_out,
                      int B, int H, int W,
                      const float* g, const float* dg, int r_sigma, const float* k, int r_rho,
                      int criterion,
                      float* d_sr, void* workspace, size_t workspace_bytes, void* stream);

/* Gradient w.r.t. gt of the three patch losses: the reference's gather of the selected candidates is differentiable
 * in p2_cat (loss.py:136-139, :219-222, :369-371), so a gt that requires grad receives, per query, minus the SR-side
 * criterion gradient pushed through the descriptor of the SELECTED candidate, accumulated over the queries that chose
 * it, and folded back through the bicubic pyramid taps for candidates of the two coarse levels.
 *   mode: 0 = BestBuddyLoss, 1 = GramLoss, 2 = PatchwiseStructureTensorLoss (g, dg, k as for srst_pst_forward;
 *         ignored, may be NULL, for modes 0 and 1).
 *   d_gt [B,3,H,W] is overwritten.  Workspace: srst_bb_workspace_bytes (always required).  Sums over queries use
 *   atomicAdd: the result can differ in the last bits from run to run. */
int srst_patch_backward_gt(int mode, const float* sr, const float* gt, const float* gt2, const float* gt4,
                           const int64_t* idx, const float* grad_out,
                           int B, int H, int W,
                           const float* g, const float* dg, int r_sigma, const float* k, int r_rho,
                           int criterion,
                           float* d_gt, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Best-Buddy loss with an arbitrary patch geometry: BestBuddyLoss(ksize, pad, stride) of the reference
 * (loss.py:86; F.unfold(kernel_size=ksize, padding=pad, stride=stride) at loss.py:116-129) for everything
 * but the default (3, 0, 3), which the tuned entry points above serve.  1 <= ksize <= 8, 0 <= pad <= 64, 1 <= stride <= 4096
 * (srst_bbg_supported);
 * patches may overlap (stride < ksize), leave gaps (stride > ksize) and reach into the zero padding.
 *   N = ny*nx, ny = (H + 2 pad - ksize)/stride + 1 (srst_bbg_num_patches); every pyramid level must hold at least
 *   one patch (F.unfold raises otherwise): SRST_E_SHAPE.  D = 3*ksize*ksize elements per patch.
 * Every (query, candidate) pair is scored exactly (no filter); criterion / SRST_BB_DIST_L1, gt2 / gt4 and the
 * first-minimum rule as for srst_bb_forward.  Workspace: srst_bbg_workspace_bytes (always required; 0 = geometry or
 * shape not usable).
 * Backward: d_sr [B,3,H,W] sums, per pixel and in a fixed order, the criterion gradients grad_out/(B*N*D) *
 * sign(sr_patch - cand[idx]) (or 2*diff) of the patches that cover the pixel; d_gt [B,3,H,W] is the gradient through
 * the gather of the selected candidates (loss.py:136-139; atomicAdd, then the bicubic pyramid adjoint).  Either of
 * d_sr / d_gt may be NULL.
 * ------------------------------------------------------------------------------------------- */
int srst_bbg_supported(int ksize, int pad, int stride);
size_t srst_bbg_workspace_bytes(int B, int H, int W, int ksize, int pad, int stride);
long long srst_bbg_num_patches(int H, int W, int ksize, int pad, int stride);
int srst_bbg_forward(const float* sr, const float* gt, const float* gt2, const float* gt4,
                     int B, int H, int W, int ksize, int pad, int stride,
                     float alpha, float beta, int criterion,
                     int64_t* idx_out, float* loss_out,
                     void* workspace, size_t workspace_bytes, void* stream);
int srst_bbg_backward(const float* sr, const float* gt, const float* gt2, const float* gt4,
                      const int64_t* idx, const float* grad_out,
                      int B, int H, int W, int ksize, int pad, int stride, int criterion,
                      float* d_sr, float* d_gt,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The HR pyramid on its own (exposed for tests): out2 [B,3,H/2,W/2], out4 [B,3,H/4,W/4]. */
int srst_bb_pyramid(const float* gt, int B, int H, int W, float* out2, float* out4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRST_H_ */
