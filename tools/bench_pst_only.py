import os, sys, torch
sys.path.insert(0, os.getcwd())
from srgan_st_b200 import PatchwiseStructureTensorLoss, GramLoss
torch.manual_seed(0)
gt = torch.rand(64, 3, 192, 192, device="cuda")
x = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1).requires_grad_(True)
for cls in (PatchwiseStructureTensorLoss, GramLoss):
    m = cls()
    for _ in range(3):
        m(x, gt).backward()
    torch.cuda.synchronize()
