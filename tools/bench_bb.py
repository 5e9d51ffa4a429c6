"""Time the Best-Buddy path on the B200 (BASELINE config 4 and the 96x96 training size)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srgan_st_b200 import BestBuddyLoss, GramLoss, PatchwiseStructureTensorLoss  # noqa: E402


def run(B, H, W, pyramid, iters=10, cls=BestBuddyLoss, d=27, dist_norm="l2"):
    torch.manual_seed(0)
    gt = torch.rand(B, 3, H, W, device="cuda")
    x = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1).requires_grad_(True)
    m = cls(pyramid=pyramid, dist_norm=dist_norm)
    for _ in range(3):
        l = m(x, gt); l.backward()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        e[0].record(); l = m(x, gt); e[1].record(); l.backward(); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    tf /= iters; tb /= iters
    N = (H // 3) * (W // 3); M = N + ((H // 2) // 3) * ((W // 2) // 3) + ((H // 4) // 3) * ((W // 4) // 3)
    gflop = 4.0 * N * M * d * B / 1e9
    print(f"{cls.__name__} dist_norm={dist_norm} B={B} {H}x{W} pyramid={pyramid}: fwd {tf:.3f} ms  bwd {tb:.3f} ms  -> {B/((tf+tb)*1e-3):.0f} img/s, "
          f"{gflop/(tf*1e-3)/1e3:.1f} TFLOP/s fp32 (algorithmic 4NMd)", flush=True)


if __name__ == "__main__":
    for pyr in ("aten", "fused"):
        run(64, 192, 192, pyr)
        run(16, 96, 96, pyr)
        run(64, 96, 96, pyr)
    run(64, 192, 192, "fused", cls=GramLoss, d=9)
    run(64, 192, 192, "fused", cls=PatchwiseStructureTensorLoss, d=27)
    run(64, 192, 192, "fused", dist_norm="l1")            # every pair scored exactly: 4 FADD per term and pair
    run(64, 192, 192, "fused", cls=GramLoss, d=9, dist_norm="l1")
