"""Time the generic-radius ST path (sigma / rho beyond the compiled classes) against the default-radius kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srgan_st_b200 import StructureTensorLoss  # noqa: E402


def run(B, H, W, sigma, rho, iters=20):
    torch.manual_seed(0)
    y = torch.rand(B, 3, H, W, device="cuda")
    x = (y + 0.05 * torch.randn_like(y)).clamp(0, 1).requires_grad_(True)
    m = StructureTensorLoss(sigma=sigma, rho=rho)
    for _ in range(3):
        m(x, y).backward()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        e[0].record(); l = m(x, y); e[1].record(); l.backward(); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print(f"StructureTensorLoss sigma={sigma} rho={rho} B={B} {H}x{W}: fwd {tf / iters * 1e3:.1f} us  bwd {tb / iters * 1e3:.1f} us "
          f"(eager module call, includes host overhead)", flush=True)


if __name__ == "__main__":
    for s, r in ((0.5, 2.0), (1.0, 3.0), (1.5, 4.0), (2.0, 8.0), (4.0, 16.0)):
        run(64, 96, 96, s, r)
    for s, r in ((0.5, 2.0), (1.5, 4.0)):
        run(1, 1356, 2040, s, r)
