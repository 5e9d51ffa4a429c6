"""GramLoss (reference loss.py:146-225): oracle pinned on the reference's outputs, kernels checked
under the host emulation (CPU) and on the GPU."""
import numpy as np
import pytest

from oracle import bb_oracle as O
from tests.helpers import emu_bb, emu_lib, golden, golden_names, maxnorm_err, rel_err

GRAM_CASES = golden_names("gram_")


def test_fixture_inventory():
    assert len(GRAM_CASES) >= 3


@pytest.mark.parametrize("name", GRAM_CASES)
def test_oracle_matches_reference(name):
    z = golden(name)
    crit = str(z["criterion"])
    r = O.bb_forward_c(z["sr"], z["hr"], z["hr2"], z["hr4"], float(z["alpha"]), float(z["beta"]), crit, mode="gram")
    assert np.abs(O.gram_descriptors(z["sr"]) - z["p1"]).max() < 1e-6          # descriptor = reference's gram_matrix
    assert np.array_equal(r["idx"], z["ind"])
    assert rel_err(r["loss"], z["loss"]) < 1e-6
    cat = np.concatenate([O.gram_descriptors(z[k]) for k in ("hr", "hr2", "hr4")], 1)
    g = O.gram_backward(z["sr"], np.take_along_axis(cat, z["ind"][..., None], 1), crit)
    assert maxnorm_err(g, z["d_sr"]) < 1e-5


@pytest.mark.parametrize("name", GRAM_CASES)
@pytest.mark.parametrize("own_pyramid", [False, True])
def test_emulated_kernels_match_oracle_and_reference(name, own_pyramid):
    lib = emu_lib()
    z = golden(name)
    crit = str(z["criterion"])
    gt2, gt4 = (None, None) if own_pyramid else (z["hr2"], z["hr4"])
    out = emu_bb(lib, z["sr"], z["hr"], gt2, gt4, float(z["alpha"]), float(z["beta"]), 0 if crit == "l1" else 1,
                 mode="gram")
    orc = O.bb_forward_c(z["sr"], z["hr"], gt2, gt4, float(z["alpha"]), float(z["beta"]), crit, mode="gram")
    assert np.array_equal(out["idx"], orc["idx"])
    # vs the reference: equal wherever its own top-2 gap exceeds fp32 noise (gram scores are ~1e-3)
    gap = z["top2"][..., 1] - z["top2"][..., 0]
    differ = out["idx"] != z["ind"]
    assert not (differ & (gap > 1e-6 * np.maximum(z["top2"][..., 1], 1e-6) + 1e-9)).any()
    assert rel_err(out["loss"], z["loss"]) < 1e-5
    assert maxnorm_err(out["d_sr"], z["d_sr"]) < 1e-4 or differ.any()


def test_emulated_gram_ragged_shape():
    lib = emu_lib()
    rng = np.random.default_rng(9)
    sr = rng.random((1, 3, 26, 31), dtype=np.float32)
    gt = rng.random((1, 3, 26, 31), dtype=np.float32)
    out = emu_bb(lib, sr, gt, mode="gram")
    orc = O.bb_forward_c(sr, gt, mode="gram")
    assert np.array_equal(out["idx"], orc["idx"]) and rel_err(out["loss"], orc["loss"]) < 1e-6
    assert np.all(out["d_sr"][:, :, 24:, :] == 0) and np.all(out["d_sr"][:, :, :, 30:] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GRAM_CASES)
@pytest.mark.parametrize("pyramid", ["aten", "fused"])
def test_gpu_matches_reference_golden(name, pyramid):
    import torch
    from srgan_st_b200 import GramLoss
    z = golden(name)
    crit = str(z["criterion"])
    x = torch.from_numpy(z["sr"]).cuda().requires_grad_(True)
    y = torch.from_numpy(z["hr"]).cuda()
    m = GramLoss(alpha=float(z["alpha"]), beta=float(z["beta"]), criterion=crit, pyramid=pyramid)
    loss = m(x, y)
    loss.backward()
    idx = m.last_indices.cpu().numpy()
    gap = z["top2"][..., 1] - z["top2"][..., 0]
    differ = idx != z["ind"]
    assert not (differ & (gap > 1e-6 * np.maximum(z["top2"][..., 1], 1e-6) + 1e-9)).any()
    assert rel_err(loss.item(), z["loss"]) < 1e-5
    if not differ.any():
        assert maxnorm_err(x.grad.cpu().numpy(), z["d_sr"]) < 1e-4


@pytest.mark.gpu
def test_gpu_bit_exact_vs_c_oracle_and_properties():
    import torch
    from srgan_st_b200 import GramLoss
    rng = np.random.default_rng(5)
    sr = rng.random((2, 3, 96, 96), dtype=np.float32)
    gt = rng.random((2, 3, 96, 96), dtype=np.float32)
    m = GramLoss(pyramid="fused")
    x = torch.from_numpy(sr).cuda().requires_grad_(True)
    loss = m(x, torch.from_numpy(gt).cuda())
    loss.backward()
    orc = O.bb_forward_c(sr, gt, mode="gram")
    assert np.array_equal(m.last_indices.cpu().numpy(), orc["idx"])
    assert rel_err(loss.item(), orc["loss"]) < 1e-5
    cat = np.concatenate([O.gram_descriptors(t) for t in (gt, *O.pyramid_c(gt))], 1)
    g = O.gram_backward(sr, np.take_along_axis(cat, orc["idx"][..., None], 1))
    assert maxnorm_err(x.grad.cpu().numpy(), g) < 1e-4
    same = torch.from_numpy(gt).cuda()
    assert GramLoss(pyramid="fused")(same.clone(), same).item() == 0.0
