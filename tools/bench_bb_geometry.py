"""Time BestBuddyLoss with non-default patch geometries (the exact all-pairs path, bb_generic.cuh) on the B200."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srgan_st_b200 import BestBuddyLoss, _cabi  # noqa: E402


def run(B, H, W, k, p, s, force_generic=False, iters=5):
    torch.manual_seed(0)
    gt = torch.rand(B, 3, H, W, device="cuda")
    x = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1).requires_grad_(True)
    m = BestBuddyLoss(ksize=k, pad=p, stride=s)
    if force_generic:
        m._geom = (k, p, s)
    for _ in range(2):
        l = m(x, gt); l.backward()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        e[0].record(); l = m(x, gt); e[1].record(); l.backward(); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    tf /= iters; tb /= iters
    lib = _cabi.lib()
    N = lib.srst_bbg_num_patches(H, W, k, p, s)
    M = N + lib.srst_bbg_num_patches(H // 2, W // 2, k, p, s) * 0
    n = lambda h, w: ((h + 2 * p - k) // s + 1) * ((w + 2 * p - k) // s + 1)
    M = n(H, W) + n(H // 2, W // 2) + n(H // 4, W // 4)
    gflop = 4.0 * N * M * 3 * k * k * B / 1e9
    path = "generic" if m._geom is not None else "tuned"
    print(f"ksize={k} pad={p} stride={s} ({path}) B={B} {H}x{W}: N={N} M={M} D={3*k*k}  fwd {tf:.3f} ms  bwd {tb:.3f} ms  "
          f"{gflop/(tf*1e-3)/1e3:.2f} TFLOP/s fp32 (algorithmic 4NMD)", flush=True)


if __name__ == "__main__":
    run(64, 192, 192, 3, 0, 3)                        # the tuned search (filter + exact re-scoring)
    run(64, 192, 192, 3, 0, 3, force_generic=True)    # same geometry through the exact all-pairs kernels
    run(64, 192, 192, 4, 0, 4)
    run(64, 192, 192, 6, 0, 6)
    run(64, 96, 96, 4, 1, 2)
    run(16, 96, 96, 3, 1, 1)
