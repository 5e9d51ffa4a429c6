"""Gradient w.r.t. gt of the three patch losses (reference loss.py:136-139, :219-222, :369-371: the gather of the selected
candidates is differentiable in p2_cat).  Fixtures `*_dgt.npz` hold the reference's own gt.grad on the inputs of the
fixture of the same name.  CPU: the float64 oracle and the emulated kernels; GPU: the nn.Module path."""
import numpy as np
import pytest

from oracle import bb_oracle as O
from tests.helpers import emu_lib, emu_patch_gt, golden, golden_names, maxnorm_err, rel_err

CASES = [n[:-4] for n in golden_names("", "_dgt")]


def _mode(name):
    return {"bb": "patch", "gram": "gram", "pst": "pst"}[name.split("_")[0]]


def _taps(z, mode):
    return (z["g"], z["dg"], z["k"]) if mode == "pst" else None


def test_fixture_inventory():
    assert len(CASES) >= 7 and {_mode(n) for n in CASES} == {"patch", "gram", "pst"}


@pytest.mark.parametrize("name", CASES)
def test_oracle_gt_gradient_matches_reference(name):
    z, zg = golden(name), golden(name + "_dgt")
    mode = _mode(name)
    assert rel_err(zg["loss"], z["loss"]) < 1e-6 and np.array_equal(zg["d_sr"], z["d_sr"])   # same run, same inputs
    d = O.patch_backward_gt(z["sr"], z["hr"], z["hr2"], z["hr4"], z["ind"], mode, _taps(z, mode), str(z["criterion"]))
    assert maxnorm_err(d, zg["d_gt"]) < 2e-5
    assert np.abs(zg["d_gt"]).max() > 0


@pytest.mark.parametrize("name", CASES)
def test_emulated_kernels_gt_gradient(name):
    lib = emu_lib()
    z, zg = golden(name), golden(name + "_dgt")
    mode = _mode(name)
    crit = 0 if str(z["criterion"]) == "l1" else 1
    for pyr in ((z["hr2"], z["hr4"]), (None, None)):   # the reference's pyramid, and the library's own
        d = emu_patch_gt(lib, z["sr"], z["hr"], z["ind"], pyr[0], pyr[1], crit, 1.0, mode, _taps(z, mode))
        assert not np.isnan(d).any()
        assert maxnorm_err(d, zg["d_gt"]) < 1e-4
    d2 = emu_patch_gt(lib, z["sr"], z["hr"], z["ind"], z["hr2"], z["hr4"], crit, -2.0, mode, _taps(z, mode))
    assert np.allclose(d2, -2.0 * d, rtol=1e-4, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("pyramid", ["fused", "aten"])
@pytest.mark.parametrize("name", CASES)
def test_module_gt_gradient_matches_reference(name, pyramid):
    import torch
    import srgan_st_b200 as S
    z, zg = golden(name), golden(name + "_dgt")
    mode = _mode(name)
    kw = dict(alpha=float(z["alpha"]), beta=float(z["beta"]), criterion=str(z["criterion"]), pyramid=pyramid)
    if mode == "pst":
        kw.update(sigma=float(z["sigma"]), rho=float(z["rho"]))
    m = {"patch": S.BestBuddyLoss, "gram": S.GramLoss, "pst": S.PatchwiseStructureTensorLoss}[mode](**kw)
    x = torch.from_numpy(z["sr"]).cuda().requires_grad_(True)
    y = torch.from_numpy(z["hr"]).cuda().requires_grad_(True)
    loss = m(x, y)
    (loss * 0.5).backward()
    assert rel_err(loss.item(), z["loss"]) < 1e-5
    assert np.array_equal(m.last_indices.cpu().numpy(), z["ind"])
    assert maxnorm_err(2.0 * x.grad.cpu().numpy(), z["d_sr"]) < 1e-4
    assert maxnorm_err(2.0 * y.grad.cpu().numpy(), zg["d_gt"]) < 1e-4
    # gt alone requires grad (e.g. when the "ground truth" is itself a network output)
    x2 = torch.from_numpy(z["sr"]).cuda()
    y2 = torch.from_numpy(z["hr"]).cuda().requires_grad_(True)
    m(x2, y2).backward()
    assert maxnorm_err(y2.grad.cpu().numpy(), zg["d_gt"]) < 1e-4


@pytest.mark.gpu
def test_gt_gradient_full_size_against_oracle():
    """BASELINE configs[3] crop size, batch 2: gt.grad against the float64 oracle evaluated on the kernel's own indices."""
    import torch
    import srgan_st_b200 as S
    torch.manual_seed(3)
    gt = torch.rand(2, 3, 192, 192, device="cuda")
    x = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)
    y = gt.clone().requires_grad_(True)
    m = S.BestBuddyLoss(pyramid="aten")
    m(x, y).backward()
    hr2 = torch.nn.functional.interpolate(gt, scale_factor=0.5, mode="bicubic", align_corners=False)
    hr4 = torch.nn.functional.interpolate(gt, scale_factor=0.25, mode="bicubic", align_corners=False)
    d = O.patch_backward_gt(x.cpu().numpy(), gt.cpu().numpy(), hr2.cpu().numpy(), hr4.cpu().numpy(),
                            m.last_indices.cpu().numpy(), "patch", None, "l1")
    assert maxnorm_err(y.grad.cpu().numpy(), d) < 1e-4
