"""Fused ST + "Pixel" MSE criterion: one pass per direction must give exactly the two
criteria the reference's loops evaluate separately (config.py:71-93, warmup.py:88-96): the ST golden
fixtures supply the ST term and its gradient (outputs of the reference itself), MSELoss is restated in
float64 numpy (mean((sr-hr)^2), gradient 2(sr-hr)/n)."""
import numpy as np
import pytest

from oracle import st_oracle as O
from tests.helpers import emu_lib, emu_stpx, golden, maxnorm_err, rel_err

CASES = ["st_rand_2x24x36", "st_srlike_2x40x52", "st_rand_ragged_1x37x53"]


def _mse(sr, hr):
    d = np.asarray(sr, np.float64) - np.asarray(hr, np.float64)
    return float((d * d).mean()), 2.0 * d / d.size


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cfg", [-1, 0, 1])
def test_emulated_fused_terms_and_gradient(name, cfg):
    lib = emu_lib()
    assert lib.srst_st_force_cfg(cfg, cfg) == 0   # the fused term is compiled into the two default tile shapes
    z = golden(name)
    taps = (z["g"], z["dg"], z["k"])
    w_st, w_px = 1.0 / 3.0, 1.7
    out = emu_stpx(lib, z["sr"], z["hr"], taps, grad_st=w_st, grad_px=w_px)
    mse, dmse = _mse(z["sr"], z["hr"])
    assert rel_err(out["st"], z["loss"]) < 1e-5            # the reference's own ST loss
    assert rel_err(out["px"], mse) < 1e-5
    ref = O.st_loss(z["sr"], z["hr"], taps=taps)
    want = w_st * ref["d_sr"] + w_px * dmse
    assert maxnorm_err(out["d_sr"], want) < 1e-4
    assert np.all(out["ws"] == 0), "workspace must be left zeroed (both partial arrays)"
    lib.srst_st_force_cfg(-1, -1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_the_two_reference_criteria(name):
    import torch
    from srgan_st_b200 import StructureTensorLoss, StructureTensorPixelLoss
    z = golden(name)
    x = torch.from_numpy(z["sr"]).cuda().requires_grad_(True)
    y = torch.from_numpy(z["hr"]).cuda()
    w_st, w_px = 1.0 / 3.0, 1.0
    m = StructureTensorPixelLoss(st_weight=w_st, pixel_weight=w_px)
    loss = m(x, y)
    loss.backward()
    # the two criteria evaluated separately, as the reference loop does
    x2 = torch.from_numpy(z["sr"]).cuda().requires_grad_(True)
    sep = w_st * StructureTensorLoss()(x2, y) + w_px * torch.nn.MSELoss()(x2, y)
    sep.backward()
    assert rel_err(m.last_terms[0].item(), z["loss"]) < 1e-5
    assert rel_err(m.last_terms[1].item(), _mse(z["sr"], z["hr"])[0]) < 1e-5
    assert rel_err(loss.item(), sep.item()) < 1e-6
    assert maxnorm_err(x.grad.cpu().numpy(), x2.grad.cpu().numpy()) < 1e-5
    want = w_st * z["d_sr"].astype(np.float64) + w_px * _mse(z["sr"], z["hr"])[1]
    ref64 = O.st_loss(z["sr"], z["hr"], taps=(z["g"], z["dg"], z["k"]))
    tol = 1e-4 + maxnorm_err(z["d_sr"], ref64["d_sr"])     # the reference's own fp32 noise
    assert maxnorm_err(x.grad.cpu().numpy(), want) < tol


@pytest.mark.gpu
def test_gpu_full_size_warmup_batch():
    """BASELINE configs[1] shape (batch 64 of 96x96): fused == separate, and the ST-only entry point
    is unaffected by the extra partial array in the shared workspace."""
    import torch
    from srgan_st_b200 import StructureTensorLoss, StructureTensorPixelLoss
    torch.manual_seed(3)
    y = torch.rand(64, 3, 96, 96, device="cuda")
    x = (y + 0.05 * torch.randn_like(y)).clamp(0, 1).requires_grad_(True)
    fused = StructureTensorPixelLoss(st_weight=0.5, pixel_weight=2.0)
    lf = fused(x, y); lf.backward()
    gf = x.grad.clone(); x.grad = None
    ls = 0.5 * StructureTensorLoss()(x, y) + 2.0 * torch.nn.functional.mse_loss(x, y)
    ls.backward()
    assert rel_err(lf.item(), ls.item()) < 1e-6
    assert maxnorm_err(gf.cpu().numpy(), x.grad.cpu().numpy()) < 1e-5
    again = StructureTensorLoss()(x.detach(), y).item()
    assert rel_err(again, fused.last_terms[0].item()) < 1e-6
