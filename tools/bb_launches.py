"""One forward + backward of each patch loss at BASELINE config 4 (batch 64 x 192x192) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bb_launches.csv python tools/bb_launches.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srgan_st_b200 import BestBuddyLoss, GramLoss, PatchwiseStructureTensorLoss  # noqa: E402

torch.manual_seed(0)
gt = torch.rand(64, 3, 192, 192, device="cuda")
x = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1).requires_grad_(True)
for cls in (BestBuddyLoss, GramLoss, PatchwiseStructureTensorLoss):
    m = cls()
    for _ in range(2):
        m(x, gt).backward()
    torch.cuda.synchronize()
