"""Per-role cycle accounting of the streaming forward kernel (CTA 0): SRST_ST_STREAM_DEBUG=1."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SRST_ST_STREAM"] = "1"; os.environ["SRST_ST_STREAM_DEBUG"] = "1"
from srgan_st_b200 import _cabi, taps as T
lib = _cabi.lib(); g, dg = T.gaussian_taps(0.5); k, _ = T.gaussian_taps(2.0)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
for (B, H, W) in [(1, 1356, 2040), (64, 96, 96)]:
    sr = torch.rand(B, 3, H, W, device="cuda"); hr = torch.rand(B, 3, H, W, device="cuda")
    ds = torch.empty_like(sr); loss = torch.zeros((), device="cuda")
    ws = torch.zeros(max(lib.srst_st_workspace_bytes(B, H, W), 8192), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        _cabi.check(lib.srst_st_forward(vp(sr), vp(hr), B, H, W, T.as_c(g), T.as_c(dg), 2, T.as_c(k), 8, 1, 1e-12, vp(loss), vp(ds), None, None, None, vp(ws), ws.numel(), None), "fwd")
    torch.cuda.synchronize()
    d = ws[2048:2048 + 8 * 64].view(torch.int64).cpu().tolist()
    roles = ["HC"] * 8 + ["VS"] * 5 + ["GR"] * 4 + ["LG"] * 4
    print(f"B={B} {H}x{W}")
    for w, r in enumerate(roles):
        print(f"  warp {w:2d} {r}: work {d[2*w]:8d}  wait {d[2*w+1]:8d}  ({100*d[2*w]/max(d[2*w]+d[2*w+1],1):.0f}% busy)")
    ws[2048:2048 + 8 * 64] = 0
