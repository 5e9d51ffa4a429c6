"""Phase time line of the ST forward kernel (needs tools/build_timing.sh's libsrst_timing.so): thread 0 of every CTA
stamps %globaltimer at each phase boundary; printed: per phase the median / max duration over CTAs, and the CTA
start / end spread, for one launch on L2-cold inputs.
    python tools/phase_timing.py [B H W [fwd_cfg]]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srgan_st_b200 import _cabi, taps as T  # noqa: E402

lib = ctypes.CDLL(os.path.join(os.path.dirname(_cabi.LIB_PATH), "libsrst_timing.so"))
for name, (res, args) in _cabi.SIGNATURES.items():
    fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
lib.srst_debug_set_timing.argtypes = [ctypes.c_void_p]
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 96, 96)
cfg = int(sys.argv[4]) if len(sys.argv) >= 5 else -1
lib.srst_st_force_cfg(cfg, -1)
dev = torch.device("cuda:0")
g, dg = T.gaussian_taps(0.5); k, _ = T.gaussian_taps(2.0)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
NMAX = 1 << 16
stamps = torch.zeros(NMAX * 32, dtype=torch.int64, device=dev)
pairs = [(torch.rand(B, 3, H, W, device=dev), torch.rand(B, 3, H, W, device=dev)) for _ in range(3)]
ds = torch.empty(B, 3, H, W, device=dev); ixy = torch.empty(lib.srst_st_ixy_floats(B, H, W), device=dev)
loss = torch.zeros((), device=dev)
ws = torch.zeros(max(lib.srst_st_workspace_bytes(B, H, W), 4096), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def fwd(i):
    sr, hr = pairs[i % 3]
    rc = lib.srst_st_forward(vp(sr), vp(hr), B, H, W, T.as_c(g), T.as_c(dg), 2, T.as_c(k), 8, 1, 1e-12, vp(loss), vp(ds), None,
                             vp(ixy), None, vp(ws), ws.numel(), s)
    assert rc == 0, rc


for i in range(3):
    fwd(i)
torch.cuda.synchronize()
flush.zero_()                      # evict the inputs from L2
torch.cuda.synchronize()
lib.srst_debug_set_timing(ctypes.c_void_p(stamps.data_ptr()))
fwd(0)
torch.cuda.synchronize()
lib.srst_debug_set_timing(None)
st = stamps.cpu().numpy().reshape(NMAX, 32)
n = int((st[:, 15] != 0).sum())
st = st[:n].astype(np.float64)
t0 = st[:, 15].min()
names = ["launch->SR.A", "SR.A load", "SR.B grad", "SR.C vert", "SR.D horiz", "HR.A load", "HR.B grad", "HR.C vert", "HR.D horiz",
         "chain+stores (thread 0)", "wait for the CTA's last warp"]
edges = [15, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
print(f"B={B} {H}x{W} cfg={cfg}: {n} CTAs; kernel span {(st[:, 10].max() - t0) / 1e3:.2f} us; CTA start spread "
      f"{(st[:, 15].max() - t0) / 1e3:.2f} us; CTA lifetime median {np.median(st[:, 10] - st[:, 15]) / 1e3:.2f} us, "
      f"max {(st[:, 10] - st[:, 15]).max() / 1e3:.2f} us")
for nm, a, b in zip(names, edges[:-1], edges[1:]):
    d = (st[:, b] - st[:, a]) / 1e3
    print(f"  {nm:32s} median {np.median(d):6.2f} us   p90 {np.percentile(d, 90):6.2f}   max {d.max():6.2f}")
