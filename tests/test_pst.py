"""PatchwiseStructureTensorLoss (reference loss.py:292-375): oracle pinned on the reference's outputs,
kernels checked under the host emulation (CPU) and on the GPU."""
import numpy as np
import pytest

from oracle import bb_oracle as O
from tests.helpers import emu_bb, emu_lib, golden, golden_names, maxnorm_err, rel_err

PST_CASES = golden_names("pst_")


def _taps(z):
    return (z["g"], z["dg"], z["k"])


def _near_tie_ok(idx, z):
    """Indices may differ from the reference's only where its own top-2 gap is inside fp32 noise."""
    gap = z["top2"][..., 1] - z["top2"][..., 0]
    differ = idx != z["ind"]
    return not (differ & (gap > 1e-5 * np.maximum(z["top2"][..., 1], 1e-6) + 1e-9)).any(), differ


def test_fixture_inventory():
    assert len(PST_CASES) >= 3


@pytest.mark.parametrize("name", PST_CASES)
def test_oracle_matches_reference(name):
    z = golden(name)
    crit, taps = str(z["criterion"]), _taps(z)
    assert np.abs(O.pst_descriptors(z["sr"], taps) - z["p1"]).max() < 2e-6     # = reference's vmap(vmap(s_norm))
    assert np.abs(O.describe_c(z["sr"], "pst", taps) - z["p1"]).max() < 2e-6   # fp32 C restatement too
    r = O.bb_forward_c(z["sr"], z["hr"], z["hr2"], z["hr4"], float(z["alpha"]), float(z["beta"]), crit, mode="pst",
                       taps=taps)
    assert np.array_equal(r["idx"], z["ind"])
    assert rel_err(r["loss"], z["loss"]) < 1e-6
    cat = np.concatenate([O.pst_descriptors(z[k], taps) for k in ("hr", "hr2", "hr4")], 1)
    g = O.pst_backward(z["sr"], np.take_along_axis(cat, z["ind"][..., None], 1), taps, crit)
    assert maxnorm_err(g, z["d_sr"]) < 1e-5


def test_oracle_backward_is_the_gradient_of_its_forward():
    """Finite-difference check of the hand-derived adjoint (float64)."""
    rng = np.random.default_rng(0)
    sr = rng.random((1, 3, 12, 12))
    z = golden(PST_CASES[0])
    taps = _taps(z)
    sel = rng.random((1, 16, 27))
    f = lambda x: ((O.pst_descriptors(x, taps) - sel) ** 2).mean()
    g = O.pst_backward(sr, sel, taps, "l2")
    for (c, y, x) in [(0, 0, 0), (1, 4, 7), (2, 11, 11), (0, 5, 5)]:
        e = np.zeros_like(sr)
        e[0, c, y, x] = 1e-6
        fd = (f(sr + e) - f(sr - e)) / 2e-6
        assert abs(fd - g[0, c, y, x]) < 1e-6 * max(1.0, abs(fd) * 1e3)


@pytest.mark.parametrize("name", PST_CASES)
@pytest.mark.parametrize("own_pyramid", [False, True])
def test_emulated_kernels_match_oracle_and_reference(name, own_pyramid):
    lib = emu_lib()
    z = golden(name)
    crit, taps = str(z["criterion"]), _taps(z)
    gt2, gt4 = (None, None) if own_pyramid else (z["hr2"], z["hr4"])
    out = emu_bb(lib, z["sr"], z["hr"], gt2, gt4, float(z["alpha"]), float(z["beta"]), 0 if crit == "l1" else 1,
                 mode="pst", taps=taps)
    orc = O.bb_forward_c(z["sr"], z["hr"], gt2, gt4, float(z["alpha"]), float(z["beta"]), crit, mode="pst", taps=taps)
    assert np.array_equal(out["idx"], orc["idx"]), "indices must be bit-exact vs the C oracle"
    ok, differ = _near_tie_ok(out["idx"], z)
    assert ok
    assert rel_err(out["loss"], z["loss"]) < 1e-5
    assert maxnorm_err(out["d_sr"], z["d_sr"]) < 1e-4 or differ.any()


def test_emulated_pst_ragged_shape():
    lib = emu_lib()
    rng = np.random.default_rng(9)
    z = golden(PST_CASES[0])
    sr = rng.random((1, 3, 26, 31), dtype=np.float32)
    gt = rng.random((1, 3, 26, 31), dtype=np.float32)
    out = emu_bb(lib, sr, gt, mode="pst", taps=_taps(z))
    orc = O.bb_forward_c(sr, gt, mode="pst", taps=_taps(z))
    assert np.array_equal(out["idx"], orc["idx"]) and rel_err(out["loss"], orc["loss"]) < 1e-6
    assert np.all(out["d_sr"][:, :, 24:, :] == 0) and np.all(out["d_sr"][:, :, :, 30:] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PST_CASES)
@pytest.mark.parametrize("pyramid", ["aten", "fused"])
def test_gpu_matches_reference_golden(name, pyramid):
    import torch
    from srgan_st_b200 import PatchwiseStructureTensorLoss
    z = golden(name)
    crit = str(z["criterion"])
    x = torch.from_numpy(z["sr"]).cuda().requires_grad_(True)
    y = torch.from_numpy(z["hr"]).cuda()
    m = PatchwiseStructureTensorLoss(sigma=float(z["sigma"]), rho=float(z["rho"]), alpha=float(z["alpha"]),
                                     beta=float(z["beta"]), criterion=crit, pyramid=pyramid)
    loss = m(x, y)
    loss.backward()
    ok, differ = _near_tie_ok(m.last_indices.cpu().numpy(), z)
    assert ok
    assert rel_err(loss.item(), z["loss"]) < 1e-5
    if not differ.any():
        assert maxnorm_err(x.grad.cpu().numpy(), z["d_sr"]) < 1e-4


@pytest.mark.gpu
def test_gpu_bit_exact_vs_c_oracle_and_properties():
    import torch
    from srgan_st_b200 import PatchwiseStructureTensorLoss, taps as T
    rng = np.random.default_rng(5)
    sr = rng.random((2, 3, 96, 96), dtype=np.float32)
    gt = rng.random((2, 3, 96, 96), dtype=np.float32)
    g, dg = T.gaussian_taps(0.5)
    k, _ = T.gaussian_taps(2.0)
    taps = (np.asarray(g, np.float32), np.asarray(dg, np.float32), np.asarray(k, np.float32))
    m = PatchwiseStructureTensorLoss(pyramid="fused")
    x = torch.from_numpy(sr).cuda().requires_grad_(True)
    loss = m(x, torch.from_numpy(gt).cuda())
    loss.backward()
    orc = O.bb_forward_c(sr, gt, mode="pst", taps=taps)
    assert np.array_equal(m.last_indices.cpu().numpy(), orc["idx"])
    assert rel_err(loss.item(), orc["loss"]) < 1e-5
    cat = np.concatenate([O.pst_descriptors(t, taps) for t in (gt, *O.pyramid_c(gt))], 1)
    gr = O.pst_backward(sr, np.take_along_axis(cat, orc["idx"][..., None], 1), taps)
    assert maxnorm_err(x.grad.cpu().numpy(), gr) < 1e-4
    same = torch.from_numpy(gt).cuda()
    assert PatchwiseStructureTensorLoss(pyramid="fused")(same.clone(), same).item() == 0.0


@pytest.mark.gpu
def test_gpu_constructor_contract():
    from srgan_st_b200 import PatchwiseStructureTensorLoss
    m = PatchwiseStructureTensorLoss()
    assert (m.sigma, m.rho, m.alpha, m.beta, m.ksize, m.dist_norm) == (0.5, 2, 1.0, 1.0, 3, "l2")
    with pytest.raises(NotImplementedError):
        PatchwiseStructureTensorLoss(criterion="huber")   # loss.py:323
    with pytest.raises(NotImplementedError):
        PatchwiseStructureTensorLoss(dist_norm="cosine")  # utils.py:189
