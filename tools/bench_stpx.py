"""Fused ST + Pixel criterion vs the two criteria evaluated separately (the reference loop's way), fwd+bwd."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srgan_st_b200 import StructureTensorLoss, StructureTensorPixelLoss  # noqa: E402


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for B, H, W in ((64, 96, 96), (1, 1356, 2040)):
    torch.manual_seed(0)
    y = torch.rand(B, 3, H, W, device="cuda")
    x = (y + 0.05 * torch.randn_like(y)).clamp(0, 1).requires_grad_(True)
    fused = StructureTensorPixelLoss(st_weight=1 / 3, pixel_weight=1.0)
    st, mse = StructureTensorLoss(), torch.nn.MSELoss()

    def run_fused():
        x.grad = None
        fused(x, y).backward()

    def run_sep():
        x.grad = None
        (st(x, y) / 3 + mse(x, y)).backward()

    # whole-step CUDA graphs remove the Python launch overhead from the comparison
    def graphed(fn):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        return g.replay

    tf, ts = timeit(run_fused), timeit(run_sep)
    gf, gs = timeit(graphed(run_fused)), timeit(graphed(run_sep))
    print(f"B={B} {H}x{W}: eager fused {tf:.1f} us vs separate {ts:.1f} us | graphed fused {gf:.1f} us vs separate {gs:.1f} us",
          flush=True)
