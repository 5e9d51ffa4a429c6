"""One forward of BestBuddyLoss(ksize, pad, stride) for an ncu capture of bbg_search_kernel:
    ncu --set full --clock-control none --import-source on -k regex:bbg_search -c 1 -o gpurun_out/bbg python tools/bbg_launch.py 4 0 4"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srgan_st_b200 import BestBuddyLoss  # noqa: E402

k, p, s = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 0, 4)
torch.manual_seed(0)
gt = torch.rand(64, 3, 192, 192, device="cuda")
x = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)
m = BestBuddyLoss(ksize=k, pad=p, stride=s)
for _ in range(2):
    l = m(x, gt)
torch.cuda.synchronize()
print(float(l))
