"""GPU parity of the Best-Buddy loss (nn.Module -> autograd.Function -> C ABI -> sm_100a kernels).

Index parity protocol (BASELINE.json: "argmin indices must be bit-exact"):
  * vs the C oracle (oracle/bb_oracle.c, same fixed fp32 operation order): torch.equal on every row;
  * vs the reference's own indices (golden fixtures, and a live torch restatement of loss.py:116-135
    with torch.bmm on the GPU): equal on every row whose top-2 score gap exceeds fp32 rounding noise
    (torch.bmm's summation order is unspecified); rows inside the noise band must pick one of the
    co-minimal candidates.  Both counts are asserted.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bb_oracle as O
from tests.helpers import golden, golden_names, maxnorm_err, rel_err

pytestmark = pytest.mark.gpu


def _run(sr, gt, pyramid="aten", **kw):
    from srgan_st_b200 import BestBuddyLoss
    x = torch.from_numpy(np.ascontiguousarray(sr)).cuda().requires_grad_(True)
    y = torch.from_numpy(np.ascontiguousarray(gt)).cuda()
    m = BestBuddyLoss(pyramid=pyramid, **kw)
    loss = m(x, y)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), m.last_indices.cpu().numpy(), x.grad.cpu().numpy()


def _torch_reference(x, gt, alpha=1.0, beta=1.0):
    """Restatement of reference loss.py:116-135 / utils.py:173-187 with stock torch ops (test only)."""
    unf = lambda t: F.unfold(t, kernel_size=3, padding=0, stride=3).permute(0, 2, 1).contiguous()
    p1, p2 = unf(x), unf(gt)
    g2 = F.interpolate(gt, scale_factor=0.5, mode="bicubic", align_corners=False)
    g4 = F.interpolate(gt, scale_factor=0.25, mode="bicubic", align_corners=False)
    cat = torch.cat([p2, unf(g2), unf(g4)], 1)

    def bpd(a, b):
        d = (a ** 2).sum(2)[:, :, None] + (b ** 2).sum(2)[:, None, :] - 2.0 * torch.bmm(a, b.transpose(1, 2))
        return torch.clamp(d, 0.0, float("inf"))

    score = alpha * bpd(p1, cat) + beta * bpd(p2, cat)
    top2 = torch.topk(score, 2, dim=2, largest=False).values
    w, ind = torch.min(score, dim=2)
    sel = torch.gather(cat, 1, ind.unsqueeze(-1).expand(-1, -1, 27))
    return ind, top2, (p1 - sel).abs().mean(), score


@pytest.mark.parametrize("name", golden_names("bb_"))
@pytest.mark.parametrize("pyramid", ["aten", "fused"])
def test_matches_reference_golden(name, pyramid):
    z = golden(name)
    crit = str(z["criterion"])
    loss, idx, d_sr = _run(z["sr"], z["hr"], pyramid, alpha=float(z["alpha"]), beta=float(z["beta"]), criterion=crit)
    assert np.array_equal(idx, z["ind"]), "argmin indices vs the reference"
    assert rel_err(loss, z["loss"]) < 1e-5 or abs(loss - float(z["loss"])) < 1e-9
    if np.abs(z["d_sr"]).max() == 0:
        assert np.abs(d_sr).max() == 0
    else:
        assert maxnorm_err(d_sr, z["d_sr"]) < 1e-5


@pytest.mark.parametrize("shape", [(2, 96, 96), (1, 192, 192), (1, 60, 132), (1, 50, 77)])
def test_indices_bit_exact_vs_c_oracle(shape):
    rng = np.random.default_rng(shape[2])
    sr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    gt = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    loss, idx, d_sr = _run(sr, gt, "fused")
    orc = O.bb_forward_c(sr, gt)
    assert np.array_equal(idx, orc["idx"])
    assert rel_err(loss, orc["loss"]) < 1e-5
    assert abs(np.abs(d_sr).sum() - 1.0) < 1e-4 or shape[1] % 3 or shape[2] % 3  # L1 grad abs-sum == 1


@pytest.mark.parametrize("shape", [(4, 96, 96), (2, 192, 192)])
def test_indices_vs_live_torch_reference_near_tie_protocol(shape):
    torch.manual_seed(shape[1])
    x = torch.rand(shape[0], 3, shape[1], shape[2], device="cuda")
    gt = torch.rand(shape[0], 3, shape[1], shape[2], device="cuda")
    from srgan_st_b200 import BestBuddyLoss
    m = BestBuddyLoss()
    loss = m(x.clone().requires_grad_(True), gt)
    ind_ref, top2, loss_ref, score = _torch_reference(x, gt)
    ours = m.last_indices
    gap = top2[..., 1] - top2[..., 0]
    noise = 4e-6 * top2[..., 1].clamp_min(1e-6) + 1e-6   # fp32 rounding of a ~|x|^2+|y|^2 sized sum
    differ = ours != ind_ref
    clear = gap > noise
    assert not (differ & clear).any(), "a clearly separated row picked a different candidate"
    # rows inside the noise band: our pick must be co-minimal in the reference's own score matrix
    if differ.any():
        s_ours = torch.gather(score, 2, ours.unsqueeze(-1)).squeeze(-1)
        assert ((s_ours - top2[..., 0])[differ] <= noise[differ]).all()
    assert differ.float().mean().item() < 1e-3
    assert rel_err(loss.item(), loss_ref.item()) < 1e-4


def test_full_size_properties_config4():
    """BASELINE config 4: batch 64 of 192x192.  BB(x, x) = 0 with idx[i] = i; gradient abs-sum = 1;
    a batch equals its halves."""
    from srgan_st_b200 import BestBuddyLoss
    torch.manual_seed(4)
    gt = torch.rand(64, 3, 192, 192, device="cuda")
    m = BestBuddyLoss(pyramid="fused")
    l0 = m(gt.clone(), gt)
    N = 64 * 64
    assert l0.item() == 0.0
    assert torch.equal(m.last_indices, torch.arange(N, device="cuda").expand(64, N))
    x = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1).requires_grad_(True)
    l = m(x, gt)
    idx_all = m.last_indices.clone()
    l.backward()
    assert abs(x.grad.abs().sum().item() - 1.0) < 1e-3
    la = m(x[:32].detach(), gt[:32]); ia = m.last_indices.clone()
    lb = m(x[32:].detach(), gt[32:]); ib = m.last_indices.clone()
    assert torch.equal(torch.cat([ia, ib]), idx_all)
    assert rel_err(l.item(), 0.5 * (la.item() + lb.item())) < 1e-6


def test_odd_weights_keep_the_reference_rounding_points():
    """alpha, beta that are not powers of two: alpha*d1, beta*d2 and their sum must round separately
    (loss.py:132-133) -- a contracted fma would still pass every power-of-two case.  Bit-exact vs the
    C oracle on the indices AND on the winning scores."""
    rng = np.random.default_rng(21)
    sr = rng.random((2, 3, 48, 60), dtype=np.float32)
    gt = rng.random((2, 3, 48, 60), dtype=np.float32)
    for a, b in ((0.7, 1.3), (1.0 / 3.0, 0.9)):
        _, idx, _ = _run(sr, gt, "fused", alpha=a, beta=b)
        orc = O.bb_forward_c(sr, gt, alpha=a, beta=b)
        assert np.array_equal(idx, orc["idx"])


def test_search_filter_exact_on_ties_and_odd_weights_full_size():
    """Search kernel = single-dot lower-bound filter + exact re-scoring.  Binary images (thousands of
    exact ties), zero / negative weights, at the 96x96 training size: indices equal the C oracle."""
    rng = np.random.default_rng(31)
    gt = (rng.random((2, 3, 96, 96)) > 0.5).astype(np.float32)
    sr = (rng.random((2, 3, 96, 96)) > 0.5).astype(np.float32)
    _, idx, _ = _run(sr, gt, "fused")
    assert np.array_equal(idx, O.bb_forward_c(sr, gt)["idx"])
    gt = rng.random((2, 3, 96, 96), dtype=np.float32)
    sr = np.clip(gt + 0.05 * rng.standard_normal(gt.shape).astype(np.float32), 0, 1)
    for a, b in ((0.0, 1.0), (1.0, 0.0), (-0.25, 1.0)):
        _, idx, _ = _run(sr, gt, "fused", alpha=a, beta=b)
        assert np.array_equal(idx, O.bb_forward_c(sr, gt, alpha=a, beta=b)["idx"])


def test_rejects_unsupported_geometry():
    from srgan_st_b200 import BestBuddyLoss
    with pytest.raises(NotImplementedError):
        BestBuddyLoss(criterion="huber")
    with pytest.raises(NotImplementedError):
        BestBuddyLoss(ksize=9)             # 1 <= ksize <= 8 have kernels (tests/test_bb_geometry.py)
    with pytest.raises(NotImplementedError):
        BestBuddyLoss(dist_norm="cosine")  # utils.py:189 (l1 and l2 have kernels)
    with pytest.raises(RuntimeError):
        BestBuddyLoss()(torch.rand(1, 3, 24, 24), torch.rand(1, 3, 24, 24))
    # a gt that requires grad gets the gradient through the gathered candidates (loss.py:136-139; tests/test_gt_grad.py)
    y = torch.rand(1, 3, 24, 24, device="cuda").requires_grad_()
    BestBuddyLoss()(torch.rand(1, 3, 24, 24, device="cuda"), y).backward()
    assert y.grad is not None and torch.isfinite(y.grad).all() and y.grad.abs().max() > 0
