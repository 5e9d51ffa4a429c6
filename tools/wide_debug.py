"""Per-warp phase time stamps of the wide-strip ST kernels (CTA 0): SRST_ST_DEBUG=1."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SRST_ST_DEBUG"] = "1"
os.environ.setdefault("SRST_ST_FWD_CFG", "10")
from srgan_st_b200 import _cabi, taps as T
lib = _cabi.lib(); g, dg = T.gaussian_taps(0.5); k, _ = T.gaussian_taps(2.0)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
NW = int(os.environ.get("DBG_WARPS", "12"))
for (B, H, W) in [(16, 96, 96), (64, 96, 96)]:
    sr = torch.rand(B, 3, H, W, device="cuda"); hr = torch.rand(B, 3, H, W, device="cuda")
    ds = torch.empty_like(sr); loss = torch.zeros((), device="cuda"); d_sr = torch.empty_like(sr); go = torch.ones((), device="cuda")
    ws = torch.zeros(max(lib.srst_st_workspace_bytes(B, H, W), 65536), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        _cabi.check(lib.srst_st_forward(vp(sr), vp(hr), B, H, W, T.as_c(g), T.as_c(dg), 2, T.as_c(k), 8, 1, 1e-12, vp(loss), vp(ds), None, None, None, vp(ws), ws.numel(), None), "fwd")
    torch.cuda.synchronize()
    d = ws[4096:4096 + 8 * 32 * NW].view(torch.int64).cpu().view(NW, 32)
    print(f"FWD B={B} {H}x{W}: stamps relative to warp 0 slot 0 (cycles); columns = slots")
    t0 = int(d[0, 0])
    for w in range(NW):
        print(f"  w{w:2d}: " + " ".join(f"{int(x) - t0:6d}" if int(x) else "     -" for x in d[w, :24]))
    ws[4096:4096 + 8 * 32 * NW] = 0
    if os.environ.get("DBG_BWD", "0") == "1":
        for _ in range(3):
            _cabi.check(lib.srst_st_backward(vp(sr), None, vp(ds), vp(go), B, H, W, T.as_c(g), T.as_c(dg), 2, T.as_c(k), 8, vp(d_sr), None), "bwd")
        torch.cuda.synchronize()
