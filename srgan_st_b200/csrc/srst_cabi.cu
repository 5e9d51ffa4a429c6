// C ABI of libsrst.so (declared in include/srst.h).  Host-side launch logic only: validates
// arguments, picks a compiled tile configuration and enqueues the sm_100a kernels on the caller's
// stream.  No allocation, no synchronisation, no global state, no CPU fallback.
//
// The same file compiles with g++ -DSRST_EMULATE into tests/emu/_build/libsrst_emu.so, a
// test-only host emulation used to check kernel index logic in the GPU-less build container.
#include "../../include/srst.h"

#include <cstdlib>
#include <cstring>

#include <type_traits>

#include "st_kernels.cuh"
#include "st_stream.cuh"
#include "st_wide.cuh"
#include "bb_kernels.cuh"

namespace srst {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return (s && *s) ? std::atoi(s) : dflt;
}

// ---- compiled structure-tensor tile configurations ------------------------------------------
//                         TH  TW  RS CSB NP RG RK MINB
using FwdA = StFwdCfg<40, 64, 10, 4, 0, 2, 8, 2>;   // large images (DIV2K-sized validation)
using FwdB = StFwdCfg<32, 96, 8, 4, 0, 2, 8, 2>;    // 96-wide training crops: a tile spans the row
using FwdC = StFwdCfg<48, 48, 12, 4, 0, 2, 8, 2>;   // square quarter of a 96x96 crop
using FwdD = StFwdCfg<32, 64, 16, 4, 0, 2, 8, 3>;   // small footprint: three CTAs per SM
using FwdE = StFwdCfg<32, 64, 16, 4, 64, 2, 8, 2>;  // FwdD + two producer warps, double-buffered gray tile (persistent)
using FwdF = StFwdCfg<24, 64, 12, 4, 0, 2, 8, 4>;   // experiment: four small CTAs per SM
using FwdG = StFwdCfg<32, 32, 16, 4, 0, 2, 8, 4>;   // experiment: four 128-thread CTAs per SM
using FwdH = StFwdCfg<48, 64, 12, 4, 0, 2, 8, 2>;   // experiment: taller tile, two CTAs per SM
using FwdI = StFwdCfg<48, 96, 12, 4, 0, 2, 8, 1>;   // experiment: half a 96x96 crop per CTA, one 576-thread CTA per SM
using FwdJ = StFwdCfg<64, 64, 16, 4, 0, 2, 8, 1>;   // experiment: 512-thread CTA, one per SM
//                         TH  TW  RS   NT  RG RK MINB
using BwdA = StBwdCfg<24, 64, 14, 256, 2, 8, 2>;  // large images
using BwdD = StBwdCfg<28, 56, 16, 256, 2, 8, 2>;  // 16 row pairs per phase item column: no LDS bank conflicts
using BwdG = StBwdCfg<16, 96, 10, 256, 2, 8, 2, 8>;  // BwdB with 8-column horizontal-pass items
using BwdH = StBwdCfg<16, 96, 10, 288, 2, 8, 2, 8>;  // BwdG with 288 threads: the 260 gradient items fit one round
using BwdI = StBwdCfg<12, 96, 8, 224, 2, 8, 3, 8>;   // shorter strips, three CTAs per SM
using BwdF = StBwdCfg<28, 56, 16, 256, 2, 8, 2, 8>;  // BwdD with 8-column horizontal-pass items (less shared-memory traffic)

//                             TH  TW  RS   NT
using WideA = StWideFwdCfg<48, 96, 12, 384>;  // half a 96x96 training crop per CTA, one CTA per SM

constexpr int kMinFwdTH = 24, kMinFwdTW = 32;  // finest compiled forward tiling (workspace sizing)

static bool pdl_enabled_host() {
#ifdef SRST_EMULATE
  return false;
#else
  return srst::pdl_enabled();
#endif
}

static int sm_count() {
#ifdef SRST_EMULATE
  return 2;  // small persistent grid so the emulation exercises the tile loop
#else
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
#endif
}

// Opt in to > 48 KB of dynamic shared memory once per (kernel, device).
// `Tag` makes the once-flag unique per kernel instantiation (all kernels of one direction share a
// function-pointer type).
template <class Tag, class K>
static int ensure_smem(K kernel, size_t bytes) {
#ifdef SRST_EMULATE
  (void)kernel; (void)bytes;
  return 0;
#else
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev >= 0 && dev < 64 && done[dev]) return 0;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  if (dev >= 0 && dev < 64) done[dev] = true;
  return 0;
#endif
}

// Copies the (2*rs+1)- and (2*rk+1)-tap filters into the compiled radii (RG >= rs, RK >= rk),
// centred and zero-padded: a filter with extra zero taps at both ends is the same filter.
template <int RG, int RK>
static void fill_taps(StTaps<RG, RK>& t, const float* g, const float* dg, int rs, const float* k, int rk) {
  for (int i = 0; i <= 2 * RG; ++i) {
    const int j = i - (RG - rs);
    t.g[i] = (j >= 0 && j <= 2 * rs) ? g[j] : 0.f;
    t.dg[i] = (j >= 0 && j <= 2 * rs) ? dg[j] : 0.f;
  }
  for (int i = 0; i <= 2 * RK; ++i) {
    const int j = i - (RK - rk);
    t.k[i] = (j >= 0 && j <= 2 * rk) ? k[j] : 0.f;
  }
  // tap pairs (t[u], t[u-1]) for the row-pair (FFMA2) vertical passes
  for (int u = 0; u <= 2 * RG + 1; ++u) {
    t.gp[u] = make_float2(u <= 2 * RG ? t.g[u] : 0.f, u >= 1 ? t.g[u - 1] : 0.f);
    t.dgp[u] = make_float2(u <= 2 * RG ? t.dg[u] : 0.f, u >= 1 ? t.dg[u - 1] : 0.f);
  }
  for (int u = 0; u <= 2 * RK + 1; ++u) t.kp[u] = make_float2(u <= 2 * RK ? t.k[u] : 0.f, u >= 1 ? t.k[u - 1] : 0.f);
}

template <class C> struct PxTag {};  // once-flag tag of the fused-Pixel instantiation of a forward tile shape

template <class C, bool PX = false>
static int launch_st_forward(StFwdParams<C::RG, C::RK> P, void* stream) {
  P.tiles_x = (P.W + C::TW - 1) / C::TW;
  P.tiles_y = (P.H + C::TH - 1) / C::TH;
  const long long ntiles = (long long)P.B * P.tiles_x * P.tiles_y;
  if (ntiles <= 0 || ntiles > 0x7fffffffLL) return SRST_E_SHAPE;
  int e = ensure_smem<std::conditional_t<PX, PxTag<C>, C>>(st_forward_kernel<C, PX>, C::SMEM_BYTES);
  if (e) return e;
  // persistent grid: one wave of resident CTAs, each looping over tiles
  const long long slots = (long long)sm_count() * C::MINB;
  const long long nblk = (C::NP == 0 || ntiles < slots) ? ntiles : slots;
  SRST_LAUNCH_PDL((st_forward_kernel<C, PX>), dim3((unsigned)nblk), dim3(C::NT), C::SMEM_BYTES, stream, P);
  return (int)cudaGetLastError();
}

template <class C>
static int launch_st_wide_forward(StFwdParams<C::RG, C::RK> P, void* stream) {
  P.tiles_x = (P.W + C::TW - 1) / C::TW;
  P.tiles_y = (P.H + C::TH - 1) / C::TH;
  const long long ntiles = (long long)P.B * P.tiles_x * P.tiles_y;
  if (ntiles <= 0 || ntiles > 0x7fffffffLL) return SRST_E_SHAPE;
  int e = ensure_smem<C>(st_wide_forward_kernel<C>, C::SMEM_BYTES);
  if (e) return e;
  SRST_LAUNCH_PDL(st_wide_forward_kernel<C>, dim3((unsigned)ntiles), dim3(C::NT), C::SMEM_BYTES, stream, P);
  return (int)cudaGetLastError();
}

// Tensor map of a [planes][H][W] fp32 tensor with a [3][box_h][box_w] box (no swizzle, zero OOB fill).
// Returns false when TMA cannot be used (unaligned tensor, driver entry point missing): the kernel
// then stages with cp.async instead.
static bool make_plane_map(SrstTmap* map, const float* base, long long planes, int H, int W, int box_w, int box_h,
                           int box_p) {
#ifdef SRST_EMULATE
  (void)box_w; (void)box_h; (void)box_p;
  map->base = base; map->W = W; map->H = H; map->P = (int)planes;
  return true;
#else
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static bool looked_up = false;
  if (!looked_up) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
    looked_up = true;
  }
  if (!encode || !aligned16(base) || (W % 4) != 0 || box_w > 256 || box_h > 256) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_p};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
#endif
}

template <class C, bool PX = false>
static int launch_st_backward(StBwdParams<C::RG, C::RK> P, const float* gray, void* stream) {
  const bool tma_ok = P.vec4 && env_int("SRST_ST_BWD_TMA", 1) != 0;
  P.use_tma = (tma_ok && make_plane_map(&P.ds_map, P.ds, (long long)P.B * 3, P.H, P.W, C::VW, C::SH, 3)) ? 1 : 0;
  P.use_gray = (tma_ok && gray && make_plane_map(&P.gray_map, gray, (long long)P.B, P.H, P.W, C::GW, C::GH, 1)) ? 1 : 0;
  P.tiles_x = (P.W + C::TW - 1) / C::TW;
  P.tiles_y = (P.H + C::TH - 1) / C::TH;
  const long long nblk = (long long)P.B * P.tiles_x * P.tiles_y;
  if (nblk <= 0 || nblk > 0x7fffffffLL) return SRST_E_SHAPE;
  P.early_ctas = (pdl_enabled_host() && env_int("SRST_ST_BWD_EARLY", 1) != 0) ? sm_count() * C::MINB : 0;
  int e = ensure_smem<std::conditional_t<PX, PxTag<C>, C>>(st_backward_kernel<C, PX>, C::SMEM_BYTES);
  if (e) return e;
  SRST_LAUNCH_PDL((st_backward_kernel<C, PX>), dim3((unsigned)nblk), dim3(C::NT), C::SMEM_BYTES, stream, P);
  return (int)cudaGetLastError();
}

static int pick_fwd_cfg(int H, int W) {
  const int forced = env_int("SRST_ST_FWD_CFG", -1);
  if (forced >= 0 && forced <= 10) return forced;
  // measured on B200 (profiles/r01_v2_tile_sweep.log): square 48x48 tiles win on 96-wide crops,
  // the 3-CTA/SM 32x64 tile wins on large images
  if (W <= 96) return 2;
  return 3;
}
static int pick_bwd_cfg(int B, int H, int W) {
  const int forced = env_int("SRST_ST_BWD_CFG", -1);
  if (forced >= 0 && forced <= 8) return forced;
  // measured on B200 (gpurun sweeps, round 1b): on 96-wide crops full-width strips win -- 16x96 with 288
  // threads once they fill the machine (two CTAs per SM), 12x96 with three CTAs per SM for small
  // batches; the conflict-free 28x56 tile wins everywhere else
  if (W <= 96) return ((long long)B * ((H + 15) / 16) >= 2LL * sm_count()) ? 7 : 8;
  return 5;
}

}  // namespace srst

using namespace srst;

extern "C" {

int srst_version(void) { return SRST_VERSION; }

const char* srst_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case SRST_E_INVALID: return "srst: invalid argument (null pointer, non-positive size or bad enum)";
    case SRST_E_UNSUPPORTED: return "srst: filter radius / patch geometry not compiled into libsrst";
    case SRST_E_WORKSPACE: return "srst: workspace missing, misaligned or too small";
    case SRST_E_SHAPE: return "srst: image shape not usable by this entry point";
    default: break;
  }
#ifndef SRST_EMULATE
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
#endif
  return "srst: unknown error";
}

// Compiled radius classes: r_sigma is padded up to 2 or 4, r_rho up to 4, 8 or 12 (zero taps).
// (2, 8) -- the reference default sigma=0.5, rho=2.0 -- has the tuned tile shapes; the other
// classes use one generic shape each.
int srst_st_supported(int r_sigma, int r_rho) { return (r_sigma >= 1 && r_sigma <= 4 && r_rho >= 1 && r_rho <= 12) ? 1 : 0; }

static size_t st_partials_bytes(int B, int H, int W) {
  // one float per CTA of the finest compiled tiling + the ticket counter, rounded to 256 bytes
  const size_t tiles = (size_t)B * ((H + kMinFwdTH - 1) / kMinFwdTH) * ((W + kMinFwdTW - 1) / kMinFwdTW);
  return ((tiles + 4 + 1024) * sizeof(float) + 255) / 256 * 256;  // + room for one partial per persistent CTA
}

size_t srst_st_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return 2 * st_partials_bytes(B, H, W);  // ST partials + ticket | partials of the fused Pixel term
}

}  // extern "C"

namespace srst {

struct StCall {
  const float *a, *b, *grad_out;  // forward: sr, hr ; backward: img, ds, grad_out
  float *o0, *o1, *loss_out;      // forward: ds_sr, ds_hr ; backward: d_img
  float *gray0, *gray1;           // forward: gray_sr, gray_hr outputs
  const float* gray;              // backward: saved gray planes or null
  int px;                         // forward: also reduce the fused Pixel (MSE) term into loss_out[1]
  const float *px_other, *grad_px;  // backward: other image of the pair + upstream gradient of the MSE term
  int B, H, W, normalize, vec4;
  float eps;
  void* workspace;
  const float *g, *dg, *k;
  int rs, rk;
  void* stream;
};

//                            TW  LG GR VS HC (warps per role)
using StreamA = StStreamCfg<64, 4, 4, 5, 8>;

// Streaming forward: auto-selected for (2, 8) filters on 16-byte aligned tensors when the problem
// has enough rows per SM to keep the pipeline full; SRST_ST_STREAM=0/1 forces it off/on.
template <class C>
static int launch_st_stream(const StCall& c) {
  StStreamParams P;
  P.sr = c.a; P.hr = c.b; P.ds_sr = c.o0; P.ds_hr = c.o1;
  P.ticket = reinterpret_cast<unsigned int*>(c.workspace);
  P.partials = reinterpret_cast<float*>(c.workspace) + 4;
  P.loss_out = c.loss_out;
  P.B = c.B; P.H = c.H; P.W = c.W;
  P.normalize = c.normalize; P.eps = c.eps;
  P.inv_count = (float)(1.0 / ((double)c.B * c.H * c.W));
  // SRST_ST_STREAM_DEBUG=1: CTA 0 dumps per-warp (work, wait) cycles at byte 2048 of the workspace
  P.debug = env_int("SRST_ST_STREAM_DEBUG", 0) ? reinterpret_cast<long long*>(reinterpret_cast<char*>(c.workspace) + 2048) : nullptr;
  fill_taps(P.taps, c.g, c.dg, c.rs, c.k, c.rk);
  const int nsm = sm_count();
  P.nstrips = (c.W + C::TW - 1) / C::TW;
  const long long cols = (long long)c.B * P.nstrips;
  long long want = (2LL * nsm + cols - 1) / cols;           // row segments per strip for ~2 units per SM
  const long long max_segs = c.H / 32 > 0 ? c.H / 32 : 1;
  if (want > max_segs) want = max_segs;
  if (want < 1) want = 1;
  P.segh = (int)(((c.H + want - 1) / want + 15) / 16 * 16);
  P.nsegs = (c.H + P.segh - 1) / P.segh;
  const long long nunits = cols * P.nsegs;
  if (nunits > 0x7fffffffLL) return SRST_E_SHAPE;
  P.nunits = (int)nunits;
  int e = ensure_smem<C>(st_stream_forward_kernel<C>, C::SMEM_BYTES);
  if (e) return e;
  const int grid = nunits < nsm ? (int)nunits : nsm;
  SRST_LAUNCH(st_stream_forward_kernel<C>, dim3(grid), dim3(C::NT), C::SMEM_BYTES, c.stream, P);
  return (int)cudaGetLastError();
}

template <int RG, int RK>
static int st_forward_rr(const StCall& c) {
  StFwdParams<RG, RK> P;
  P.sr = c.a; P.hr = c.b; P.ds_sr = c.o0; P.ds_hr = c.o1; P.gray_sr = c.gray0; P.gray_hr = c.gray1;
  P.ticket = reinterpret_cast<unsigned int*>(c.workspace);
  P.partials = reinterpret_cast<float*>(c.workspace) + 4;
  P.px_partials = c.px ? reinterpret_cast<float*>(reinterpret_cast<char*>(c.workspace) + st_partials_bytes(c.B, c.H, c.W))
                       : nullptr;
  P.loss_out = c.loss_out;
  P.B = c.B; P.H = c.H; P.W = c.W; P.tiles_x = P.tiles_y = 0;
  P.normalize = c.normalize; P.vec4 = c.vec4; P.eps = c.eps;
  P.inv_count = (float)(1.0 / ((double)c.B * c.H * c.W));
  // SRST_ST_DEBUG=1: CTA 0 writes per-warp phase time stamps at byte 4096 of the workspace (tools/wide_debug.py)
  P.debug = env_int("SRST_ST_DEBUG", 0) ? reinterpret_cast<long long*>(reinterpret_cast<char*>(c.workspace) + 4096) : nullptr;
  fill_taps(P.taps, c.g, c.dg, c.rs, c.k, c.rk);
  if constexpr (RG == 2 && RK == 8) {
    if (c.vec4 && !c.gray0 && !c.gray1 && !c.px && env_int("SRST_ST_STREAM", 0) == 1) return launch_st_stream<StreamA>(c);
    const int cfg = pick_fwd_cfg(c.H, c.W);
    if (c.px) {  // the fused Pixel term is compiled into the two default tile shapes only
      if (cfg == 2 || (cfg != 3 && c.W <= 96)) return launch_st_forward<FwdC, true>(P, c.stream);
      return launch_st_forward<FwdD, true>(P, c.stream);
    }
    if (cfg == 10 && c.vec4 && !c.gray0 && !c.gray1) return launch_st_wide_forward<WideA>(P, c.stream);
    switch (cfg) {
      case 0: return launch_st_forward<FwdA>(P, c.stream);
      case 1: return launch_st_forward<FwdB>(P, c.stream);
      case 2: return launch_st_forward<FwdC>(P, c.stream);
      case 4: return launch_st_forward<FwdE>(P, c.stream);
      case 5: return launch_st_forward<FwdF>(P, c.stream);
      case 6: return launch_st_forward<FwdG>(P, c.stream);
      case 7: return launch_st_forward<FwdH>(P, c.stream);
      case 8: return launch_st_forward<FwdI>(P, c.stream);
      case 9: return launch_st_forward<FwdJ>(P, c.stream);
      default: return launch_st_forward<FwdD>(P, c.stream);
    }
  } else {
    using G = StFwdCfg<32, 64, 16, 4, 0, RG, RK, (RK <= 8 ? 2 : 1)>;
    return c.px ? launch_st_forward<G, true>(P, c.stream) : launch_st_forward<G>(P, c.stream);
  }
}

template <int RG, int RK>
static int st_backward_rr(const StCall& c) {
  StBwdParams<RG, RK> P;
  std::memset(&P.ds_map, 0, sizeof(P.ds_map));
  std::memset(&P.gray_map, 0, sizeof(P.gray_map));
  P.use_tma = 0; P.use_gray = 0; P.debug = nullptr;
  P.img = c.a; P.ds = c.b; P.grad_out = c.grad_out; P.d_img = c.o0;
  P.px_other = c.px_other; P.grad_px = c.grad_px;
  P.B = c.B; P.H = c.H; P.W = c.W; P.tiles_x = P.tiles_y = 0;
  P.vec4 = c.vec4;
  P.inv_count = (float)(1.0 / ((double)c.B * c.H * c.W));
  fill_taps(P.taps, c.g, c.dg, c.rs, c.k, c.rk);
  if constexpr (RG == 2 && RK == 8) {
    if (c.px_other) {  // the fused Pixel term is compiled into the two default tile shapes only
      if (c.W <= 96) return launch_st_backward<BwdH, true>(P, c.gray, c.stream);
      return launch_st_backward<BwdF, true>(P, c.gray, c.stream);
    }
    switch (pick_bwd_cfg(c.B, c.H, c.W)) {
      case 3: return launch_st_backward<BwdD>(P, c.gray, c.stream);
      case 5: return launch_st_backward<BwdF>(P, c.gray, c.stream);
      case 6: return launch_st_backward<BwdG>(P, c.gray, c.stream);
      case 7: return launch_st_backward<BwdH>(P, c.gray, c.stream);
      case 8: return launch_st_backward<BwdI>(P, c.gray, c.stream);
      default: return launch_st_backward<BwdA>(P, c.gray, c.stream);
    }
  } else {
    using G = StBwdCfg<24, 64, (24 + 2 * RG) / 2, 256, RG, RK, (RK <= 8 ? 2 : 1)>;
    return c.px_other ? launch_st_backward<G, true>(P, c.gray, c.stream) : launch_st_backward<G>(P, c.gray, c.stream);
  }
}

static int st_dispatch(const StCall& c, bool forward) {
  const int rg = c.rs <= 2 ? 2 : 4;
  const int rk = c.rk <= 4 ? 4 : (c.rk <= 8 ? 8 : 12);
#define SRST_RR(RG_, RK_) \
  if (rg == RG_ && rk == RK_) return forward ? st_forward_rr<RG_, RK_>(c) : st_backward_rr<RG_, RK_>(c);
  SRST_RR(2, 8) SRST_RR(2, 4) SRST_RR(2, 12) SRST_RR(4, 4) SRST_RR(4, 8) SRST_RR(4, 12)
#undef SRST_RR
  return SRST_E_UNSUPPORTED;
}

}  // namespace srst

extern "C" {

int srst_st_forward(const float* sr, const float* hr, int B, int H, int W, const float* g, const float* dg,
                    int r_sigma, const float* k, int r_rho, int normalize, float eps, float* loss_out,
                    float* ds_sr, float* ds_hr, float* gray_sr, float* gray_hr, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (!sr || !hr || !g || !dg || !k || !loss_out || B <= 0 || H <= 0 || W <= 0) return SRST_E_INVALID;
  if (!srst_st_supported(r_sigma, r_rho)) return SRST_E_UNSUPPORTED;
  if (!workspace || !aligned16(workspace) || workspace_bytes < srst_st_workspace_bytes(B, H, W))
    return SRST_E_WORKSPACE;
  StCall c{};
  c.a = sr; c.b = hr; c.o0 = ds_sr; c.o1 = ds_hr; c.loss_out = loss_out; c.gray0 = gray_sr; c.gray1 = gray_hr;
  c.B = B; c.H = H; c.W = W; c.normalize = normalize ? 1 : 0; c.eps = eps;
  c.vec4 = (W % 4 == 0 && aligned16(sr) && aligned16(hr) && (!ds_sr || aligned16(ds_sr)) &&
            (!ds_hr || aligned16(ds_hr)) && (!gray_sr || aligned16(gray_sr)) && (!gray_hr || aligned16(gray_hr))) ? 1 : 0;
  c.workspace = workspace; c.g = g; c.dg = dg; c.k = k; c.rs = r_sigma; c.rk = r_rho; c.stream = stream;
  return st_dispatch(c, true);
}

int srst_st_backward(const float* img, const float* gray, const float* ds, const float* grad_out, int B, int H, int W,
                     const float* g, const float* dg, int r_sigma, const float* k, int r_rho, float* d_img,
                     void* stream) {
  if (!img || !ds || !grad_out || !g || !dg || !k || !d_img || B <= 0 || H <= 0 || W <= 0) return SRST_E_INVALID;
  if (!srst_st_supported(r_sigma, r_rho)) return SRST_E_UNSUPPORTED;
  StCall c{};
  c.a = img; c.b = ds; c.grad_out = grad_out; c.o0 = d_img; c.gray = gray;
  c.B = B; c.H = H; c.W = W;
  c.vec4 = (W % 4 == 0 && aligned16(img) && aligned16(d_img) && aligned16(ds)) ? 1 : 0;
  c.g = g; c.dg = dg; c.k = k; c.rs = r_sigma; c.rk = r_rho; c.stream = stream;
  return st_dispatch(c, false);
}

int srst_stpx_forward(const float* sr, const float* hr, int B, int H, int W, const float* g, const float* dg,
                      int r_sigma, const float* k, int r_rho, int normalize, float eps, float* loss2_out, float* ds_sr,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!sr || !hr || !g || !dg || !k || !loss2_out || B <= 0 || H <= 0 || W <= 0) return SRST_E_INVALID;
  if (!srst_st_supported(r_sigma, r_rho)) return SRST_E_UNSUPPORTED;
  if (!workspace || !aligned16(workspace) || workspace_bytes < srst_st_workspace_bytes(B, H, W))
    return SRST_E_WORKSPACE;
  StCall c{};
  c.a = sr; c.b = hr; c.o0 = ds_sr; c.loss_out = loss2_out; c.px = 1;
  c.B = B; c.H = H; c.W = W; c.normalize = normalize ? 1 : 0; c.eps = eps;
  c.vec4 = (W % 4 == 0 && aligned16(sr) && aligned16(hr) && (!ds_sr || aligned16(ds_sr))) ? 1 : 0;
  c.workspace = workspace; c.g = g; c.dg = dg; c.k = k; c.rs = r_sigma; c.rk = r_rho; c.stream = stream;
  return st_dispatch(c, true);
}

int srst_stpx_backward(const float* sr, const float* hr, const float* ds, const float* grad_st, const float* grad_px,
                       int B, int H, int W, const float* g, const float* dg, int r_sigma, const float* k, int r_rho,
                       float* d_sr, void* stream) {
  if (!sr || !hr || !ds || !grad_st || !grad_px || !g || !dg || !k || !d_sr || B <= 0 || H <= 0 || W <= 0)
    return SRST_E_INVALID;
  if (!srst_st_supported(r_sigma, r_rho)) return SRST_E_UNSUPPORTED;
  StCall c{};
  c.a = sr; c.b = ds; c.grad_out = grad_st; c.o0 = d_sr; c.px_other = hr; c.grad_px = grad_px;
  c.B = B; c.H = H; c.W = W;
  c.vec4 = (W % 4 == 0 && aligned16(sr) && aligned16(hr) && aligned16(d_sr) && aligned16(ds)) ? 1 : 0;
  c.g = g; c.dg = dg; c.k = k; c.rs = r_sigma; c.rk = r_rho; c.stream = stream;
  return st_dispatch(c, false);
}

}  // extern "C"

// ---- Best-Buddy entry points ---------------------------------------------------------------
extern "C" {

size_t srst_bb_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H < 12 || W < 12) return 0;
  return bb_carve(nullptr, bb_geom(B, H, W)).total_bytes;
}

static int bb_launch_pyramid(const float* gt, const BbGeom& g, float* o2, float* o4, void* stream) {
  const size_t n = (size_t)g.B * 3 * ((size_t)g.H2 * g.W2 + (size_t)g.H4 * g.W4);
  const unsigned nblk = (unsigned)((n + 255) / 256);
  SRST_LAUNCH(bb_pyramid_kernel, dim3(nblk), dim3(256), 0, stream, gt, o2, o4, g.B * 3, g.H, g.W, g.H2, g.W2,
              g.H4, g.W4);
  return (int)cudaGetLastError();
}

int srst_bb_pyramid(const float* gt, int B, int H, int W, float* out2, float* out4, void* stream) {
  if (!gt || !out2 || !out4 || B <= 0) return SRST_E_INVALID;
  if (H < 12 || W < 12) return SRST_E_SHAPE;
  return bb_launch_pyramid(gt, bb_geom(B, H, W), out2, out4, stream);
}

}  // extern "C"

template <int MODE>
static int bb_forward_impl(const float* sr, const float* gt, const float* gt2, const float* gt4, int B, int H, int W,
                           float alpha, float beta, int criterion, int64_t* idx_out, float* loss_out, void* workspace,
                           size_t workspace_bytes, void* stream, const PstTaps& tp = PstTaps{}) {
  constexpr int D = BbDesc<MODE>::D;
  if (!sr || !gt || !idx_out || !loss_out || B <= 0) return SRST_E_INVALID;
  if (criterion != SRST_BB_L1 && criterion != SRST_BB_L2) return SRST_E_INVALID;
  if (H < 12 || W < 12 || B > 65535) return SRST_E_SHAPE;
  if ((gt2 == nullptr) != (gt4 == nullptr)) return SRST_E_INVALID;
  const BbGeom g = bb_geom(B, H, W);
  if (!workspace || !aligned16(workspace)) return SRST_E_WORKSPACE;
  const BbWorkspace w = bb_carve(workspace, g, D);
  if (workspace_bytes < w.total_bytes) return SRST_E_WORKSPACE;
  int e;
  if (!gt2) {
    if ((e = bb_launch_pyramid(gt, g, w.pyr2, w.pyr4, stream)) != 0) return e;
    gt2 = w.pyr2;
    gt4 = w.pyr4;
  }
  const int npack = g.Npad > g.Mpad ? g.Npad : g.Mpad;
  SRST_LAUNCH(bb_pack_kernel<MODE>, dim3((npack + 255) / 256, B), dim3(256), 0, stream, sr, gt, gt2, gt4, w.mats,
              w.per_image, g, tp);
  if ((e = (int)cudaGetLastError()) != 0) return e;
  SRST_LAUNCH((bb_search_kernel<D, MODE == 2>), dim3(g.Npad / BB_QT, B), dim3(BB_NT), 0, stream, w.mats, w.per_image, g, alpha,
              beta, idx_out);
  if ((e = (int)cudaGetLastError()) != 0) return e;
  const unsigned nl = (unsigned)(((size_t)B * g.N + BB_NT - 1) / BB_NT);
  SRST_LAUNCH(bb_loss_kernel<D>, dim3(nl), dim3(BB_NT), 0, stream, w.mats, w.per_image, g, idx_out, criterion,
              w.partials, w.ticket, loss_out);
  return (int)cudaGetLastError();
}

extern "C" {

int srst_bb_forward(const float* sr, const float* gt, const float* gt2, const float* gt4, int B, int H, int W,
                    float alpha, float beta, int criterion, int64_t* idx_out, float* loss_out, void* workspace,
                    size_t workspace_bytes, void* stream) {
  return bb_forward_impl<0>(sr, gt, gt2, gt4, B, H, W, alpha, beta, criterion, idx_out, loss_out, workspace,
                            workspace_bytes, stream);
}

int srst_gram_forward(const float* sr, const float* gt, const float* gt2, const float* gt4, int B, int H, int W,
                      float alpha, float beta, int criterion, int64_t* idx_out, float* loss_out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return bb_forward_impl<1>(sr, gt, gt2, gt4, B, H, W, alpha, beta, criterion, idx_out, loss_out, workspace,
                            workspace_bytes, stream);
}

}  // extern "C"

// Central five taps (offsets -2..2) of a (2r+1)-tap filter: all a 3x3 patch image can see.
static void central5(const float* t, int r, float (&o)[5]) {
  for (int i = 0; i < 5; ++i) {
    const int j = r + i - 2;
    o[i] = (j >= 0 && j <= 2 * r) ? t[j] : 0.f;
  }
}
static int pst_taps(const float* g, const float* dg, int r_sigma, const float* k, int r_rho, PstTaps& tp) {
  if (!g || !dg || !k || r_sigma < 1 || r_rho < 1 || r_sigma > 4096 || r_rho > 4096) return SRST_E_INVALID;
  central5(g, r_sigma, tp.g);
  central5(dg, r_sigma, tp.dg);
  central5(k, r_rho, tp.k);
  return 0;
}

extern "C" {

int srst_pst_forward(const float* sr, const float* gt, const float* gt2, const float* gt4, int B, int H, int W,
                     const float* g, const float* dg, int r_sigma, const float* k, int r_rho, float alpha, float beta,
                     int criterion, int64_t* idx_out, float* loss_out, void* workspace, size_t workspace_bytes,
                     void* stream) {
  PstTaps tp;
  const int e = pst_taps(g, dg, r_sigma, k, r_rho, tp);
  if (e) return e;
  return bb_forward_impl<2>(sr, gt, gt2, gt4, B, H, W, alpha, beta, criterion, idx_out, loss_out, workspace,
                            workspace_bytes, stream, tp);
}

int srst_pst_backward(const float* sr, const float* gt, const float* gt2, const float* gt4, const int64_t* idx,
                      const float* grad_out, int B, int H, int W, const float* g, const float* dg, int r_sigma,
                      const float* k, int r_rho, int criterion, float* d_sr, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (!sr || !gt || !idx || !grad_out || !d_sr || B <= 0) return SRST_E_INVALID;
  if (criterion != SRST_BB_L1 && criterion != SRST_BB_L2) return SRST_E_INVALID;
  if (H < 12 || W < 12 || B > 65535) return SRST_E_SHAPE;
  if ((gt2 == nullptr) != (gt4 == nullptr)) return SRST_E_INVALID;
  PstTaps tp;
  int e = pst_taps(g, dg, r_sigma, k, r_rho, tp);
  if (e) return e;
  const BbGeom gm = bb_geom(B, H, W);
  if (!gt2) {
    if (!workspace || !aligned16(workspace)) return SRST_E_WORKSPACE;
    const BbWorkspace w = bb_carve(workspace, gm);
    if (workspace_bytes < w.total_bytes) return SRST_E_WORKSPACE;
    if ((e = bb_launch_pyramid(gt, gm, w.pyr2, w.pyr4, stream)) != 0) return e;
    gt2 = w.pyr2;
    gt4 = w.pyr4;
  }
  if (H % 3 != 0 || W % 3 != 0) {  // pixels outside every patch keep a zero gradient
#ifdef SRST_EMULATE
    std::memset(d_sr, 0, sizeof(float) * (size_t)B * 3 * H * W);
#else
    if ((e = (int)cudaMemsetAsync(d_sr, 0, sizeof(float) * (size_t)B * 3 * H * W, (cudaStream_t)stream)) != 0) return e;
#endif
  }
  const size_t total = (size_t)B * gm.N;
  SRST_LAUNCH(pst_backward_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, stream, sr, gt, gt2, gt4, idx,
              grad_out, gm, tp, criterion, d_sr);
  return (int)cudaGetLastError();
}

int srst_gram_backward(const float* sr, const float* gt, const float* gt2, const float* gt4, const int64_t* idx,
                       const float* grad_out, int B, int H, int W, int criterion, float* d_sr, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (!sr || !gt || !idx || !grad_out || !d_sr || B <= 0) return SRST_E_INVALID;
  if (criterion != SRST_BB_L1 && criterion != SRST_BB_L2) return SRST_E_INVALID;
  if (H < 12 || W < 12 || B > 65535) return SRST_E_SHAPE;
  if ((gt2 == nullptr) != (gt4 == nullptr)) return SRST_E_INVALID;
  const BbGeom g = bb_geom(B, H, W);
  int e;
  if (!gt2) {
    if (!workspace || !aligned16(workspace)) return SRST_E_WORKSPACE;
    const BbWorkspace w = bb_carve(workspace, g);
    if (workspace_bytes < w.total_bytes) return SRST_E_WORKSPACE;
    if ((e = bb_launch_pyramid(gt, g, w.pyr2, w.pyr4, stream)) != 0) return e;
    gt2 = w.pyr2;
    gt4 = w.pyr4;
  }
  if (H % 3 != 0 || W % 3 != 0) {  // pixels outside every patch keep a zero gradient
#ifdef SRST_EMULATE
    std::memset(d_sr, 0, sizeof(float) * (size_t)B * 3 * H * W);
#else
    if ((e = (int)cudaMemsetAsync(d_sr, 0, sizeof(float) * (size_t)B * 3 * H * W, (cudaStream_t)stream)) != 0) return e;
#endif
  }
  const size_t total = (size_t)B * g.N;
  SRST_LAUNCH(gram_backward_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, sr, gt, gt2, gt4, idx,
              grad_out, g, criterion, d_sr);
  return (int)cudaGetLastError();
}

int srst_bb_backward(const float* sr, const float* gt, const float* gt2, const float* gt4, const int64_t* idx,
                     const float* grad_out, int B, int H, int W, int criterion, float* d_sr, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (!sr || !gt || !idx || !grad_out || !d_sr || B <= 0) return SRST_E_INVALID;
  if (criterion != SRST_BB_L1 && criterion != SRST_BB_L2) return SRST_E_INVALID;
  if (H < 12 || W < 12 || B > 65535) return SRST_E_SHAPE;
  if ((gt2 == nullptr) != (gt4 == nullptr)) return SRST_E_INVALID;
  const BbGeom g = bb_geom(B, H, W);
  int e;
  if (!gt2) {
    if (!workspace || !aligned16(workspace)) return SRST_E_WORKSPACE;
    const BbWorkspace w = bb_carve(workspace, g);
    if (workspace_bytes < w.total_bytes) return SRST_E_WORKSPACE;
    if ((e = bb_launch_pyramid(gt, g, w.pyr2, w.pyr4, stream)) != 0) return e;
    gt2 = w.pyr2;
    gt4 = w.pyr4;
  }
  const size_t total = (size_t)B * 3 * H * W;
  SRST_LAUNCH(bb_backward_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, sr, gt, gt2, gt4, idx,
              grad_out, g, criterion, d_sr);
  return (int)cudaGetLastError();
}

}  // extern "C"
