"""GPU parity of the fused structure-tensor loss (through the nn.Module -> autograd.Function ->
ctypes -> C ABI -> sm_100a kernels path) against the committed reference outputs, the fp64 oracle
and size-independent properties at the BASELINE.json sizes.

Tolerances (BASELINE.json north_star): loss rel 1e-5, input gradients 1e-4 (max-norm relative), fp32.
Gradient note: on inputs where the SR and HR tensors nearly coincide at some pixel the reference's
own fp32 backward is up to ~4e-4 from the fp64 truth (1/(2 sqrt(disc)) amplification, see
DESIGN.md); there the CUDA path is held to 1e-4 against the fp64 oracle, and against the
reference to 1e-4 + the reference's own distance from the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import st_oracle as O
from tests.helpers import golden, golden_names, maxnorm_err, rel_err

pytestmark = pytest.mark.gpu


def _record(key, **vals):
    import json
    import os
    from tests.helpers import ROOT
    path = os.path.join(ROOT, "gpurun_out", "r02_parity.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    data[key] = {k: float(v) for k, v in vals.items()}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)

DEFAULT_CASES = [n for n in golden_names("st_") if "s1_r25" not in n and "s15_r4" not in n and "s2_r5" not in n]
GENERIC_CASES = [("st_rand_s15_r4_1x40x52", 1.5, 4.0), ("st_srlike_s2_r5_2x33x45", 2.0, 5.0)]


def _run(sr, hr, normalize=True, want_hr=True, sigma=0.5, rho=2.0):
    from srgan_st_b200 import StructureTensorLoss
    dev = torch.device("cuda:0")
    x = torch.from_numpy(np.ascontiguousarray(sr)).to(dev).requires_grad_(True)
    y = torch.from_numpy(np.ascontiguousarray(hr)).to(dev).requires_grad_(want_hr)
    loss = StructureTensorLoss(sigma=sigma, rho=rho, normalize=normalize)(x, y)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), x.grad.cpu().numpy(), (y.grad.cpu().numpy() if want_hr else None)


@pytest.fixture(params=[(-1, -1), (0, 0), (1, 1), (2, 0), (3, 1), (4, 0), (3, 0), (4, 1)], ids=lambda p: f"fwd{p[0]}-bwd{p[1]}")
def tile_cfg(request):
    """(-1, -1) = the library's own choice; otherwise force a compiled forward (0-2 tiled, 3-4 marching) and backward
    (0-1) tile shape (srst_st_force_cfg)."""
    from srgan_st_b200 import _cabi
    lib = _cabi.lib()
    assert lib.srst_st_num_cfgs(0) == 5 and lib.srst_st_num_cfgs(1) == 2
    assert lib.srst_st_force_cfg(*request.param) == 0
    yield request.param
    lib.srst_st_force_cfg(-1, -1)


@pytest.mark.parametrize("name", DEFAULT_CASES)
def test_matches_reference_golden(name, tile_cfg):
    z = golden(name)
    norm = bool(z["normalize"])
    loss, d_sr, d_hr = _run(z["sr"], z["hr"], normalize=norm)
    ref = O.st_loss(z["sr"], z["hr"], normalize=norm, taps=(z["g"], z["dg"], z["k"]), want_hr_grad=True)
    if "same" in name:
        assert rel_err(loss, z["loss"]) < 2e-2 and abs(loss - 1.118e-6) < 2e-8
        assert np.abs(d_sr).max() < 1e-8
        return
    assert rel_err(loss, z["loss"]) < 1e-5, "loss vs reference"
    assert rel_err(loss, ref["loss"]) < 1e-5, "loss vs fp64 oracle"
    if "nonorm" in name:
        assert np.abs(d_sr).max() == 0.0
        return
    # Gradients vs the reference's own: strict 1e-4 on every fixture but st_rand_1x96x96, where the reference's fp32
    # gradient itself sits 4.2e-4 from the float64 truth (ill-conditioned pixels) and the bound is
    # widened by exactly that distance.  The measured errors go to gpurun_out/r02_parity.json.
    rec = {}
    for tag, ours, orc, refg in (("dsr", d_sr, ref["d_sr"], z["d_sr"]), ("dhr", d_hr, ref["d_hr"], z["d_hr"])):
        e_orc, e_ref, ref_orc = maxnorm_err(ours, orc), maxnorm_err(ours, refg), maxnorm_err(refg, orc)
        rec.update({f"{tag}_vs_fp64": e_orc, f"{tag}_vs_ref": e_ref, f"{tag}_ref_vs_fp64": ref_orc})
        assert e_orc < 1e-4, "grad vs fp64 oracle"
        assert e_ref < (1e-4 + ref_orc if "st_rand_1x96x96" in name else 1e-4), "grad vs reference"
    _record("golden_" + name, loss_vs_ref=rel_err(loss, z["loss"]), loss_vs_fp64=rel_err(loss, ref["loss"]), **rec)


def test_hr_without_grad_and_no_grad_mode():
    from srgan_st_b200 import StructureTensorLoss
    z = golden("st_rand_2x24x36")
    loss, d_sr, _ = _run(z["sr"], z["hr"], want_hr=False)
    assert rel_err(loss, z["loss"]) < 1e-5
    assert maxnorm_err(d_sr, z["d_sr"]) < 2e-4
    with torch.no_grad():
        x = torch.from_numpy(z["sr"]).cuda()
        y = torch.from_numpy(z["hr"]).cuda()
        l2 = StructureTensorLoss()(x, y)
    assert rel_err(l2.item(), z["loss"]) < 1e-5 and not l2.requires_grad


def test_upstream_gradient_and_determinism():
    from srgan_st_b200 import StructureTensorLoss
    z = golden("st_srlike_2x40x52")
    x = torch.from_numpy(z["sr"]).cuda().requires_grad_(True)
    y = torch.from_numpy(z["hr"]).cuda()
    m = StructureTensorLoss()
    (m(x, y) * (1.0 / 3.0)).backward()          # train.py:138-139: loss * weight
    g1 = x.grad.clone()
    x.grad = None
    l_a = m(x, y)
    l_b = m(x, y)
    assert torch.equal(l_a, l_b), "loss reduction must be deterministic"
    l_a.backward()
    assert torch.allclose(g1, x.grad / 3.0, rtol=1e-6, atol=0)
    assert maxnorm_err(x.grad.cpu().numpy(), z["d_sr"]) < 2e-4


@pytest.mark.parametrize("shape", [(16, 96, 96), (3, 100, 152), (1, 333, 517), (2, 64, 1024)])
def test_matches_oracle_on_larger_shapes(shape):
    rng = np.random.default_rng(shape[1])
    sr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    hr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    loss, d_sr, d_hr = _run(sr, hr)
    ref = O.st_loss(sr, hr, taps=None, want_hr_grad=True)
    assert rel_err(loss, ref["loss"]) < 1e-5
    assert maxnorm_err(d_sr, ref["d_sr"]) < 1e-4
    assert maxnorm_err(d_hr, ref["d_hr"]) < 1e-4


@pytest.mark.parametrize("cfg", [3, 4])
@pytest.mark.parametrize("chunk_blocks", [1, 2, 5])
@pytest.mark.parametrize("shape", [(2, 96, 96), (1, 100, 152), (1, 333, 516)])
def test_marching_forward_row_chunks(cfg, chunk_blocks, shape):
    """Forward cfgs 6 / 7 (row-marching kernel) cut into chunks of 16, 32, 80 rows: seams, warm-up block, ragged tail."""
    from srgan_st_b200 import _cabi
    lib = _cabi.lib()
    assert lib.srst_st_force_cfg(cfg, -1) == 0 and lib.srst_st_force_chunk_blocks(chunk_blocks) == 0
    try:
        rng = np.random.default_rng(shape[1] + chunk_blocks)
        sr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
        hr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
        loss, d_sr, d_hr = _run(sr, hr)
        ref = O.st_loss(sr, hr, taps=None, want_hr_grad=True)
        assert rel_err(loss, ref["loss"]) < 1e-5
        assert maxnorm_err(d_sr, ref["d_sr"]) < 1e-4 and maxnorm_err(d_hr, ref["d_hr"]) < 1e-4
    finally:
        lib.srst_st_force_cfg(-1, -1)
        lib.srst_st_force_chunk_blocks(0)


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 5, 7), (2, 17, 4), (1, 2, 130), (3, 9, 9)])
def test_tiny_and_degenerate_shapes(shape):
    """Images smaller than a tile, a single pixel, one-row strips: everything is halo."""
    rng = np.random.default_rng(sum(shape))
    sr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    hr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    loss, d_sr, d_hr = _run(sr, hr)
    ref = O.st_loss(sr, hr, want_hr_grad=True)
    assert rel_err(loss, ref["loss"]) < 1e-5
    assert maxnorm_err(d_sr, ref["d_sr"]) < 1e-4 and maxnorm_err(d_hr, ref["d_hr"]) < 1e-4


def test_full_size_properties_div2k():
    """Config 5 size (1356x2040): batch additivity (mean of means), translation of the tiling
    (crop consistency far from borders is NOT expected -- zero padding -- so use batch splits),
    gradient of a batch equals per-image gradients scaled by 1/B."""
    from srgan_st_b200 import StructureTensorLoss
    torch.manual_seed(0)
    H, W = 1356, 2040
    hr = (torch.randint(0, 256, (2, 3, H, W), device="cuda").float() / 255)
    sr = (hr + 0.05 * torch.randn_like(hr)).clamp(0, 1).requires_grad_(True)
    m = StructureTensorLoss()
    l_all = m(sr, hr)
    l_all.backward()
    g_all = sr.grad.clone()
    parts, grads = [], []
    for i in range(2):
        s = sr.detach()[i:i + 1].clone().requires_grad_(True)
        li = m(s, hr[i:i + 1])
        li.backward()
        parts.append(li.item())
        grads.append(s.grad)
    assert rel_err(l_all.item(), 0.5 * (parts[0] + parts[1])) < 1e-6
    g_parts = torch.cat(grads) / 2
    assert (g_all - g_parts).abs().max().item() <= 1e-6 * g_parts.abs().max().item() + 1e-12
    assert torch.isfinite(g_all).all()
    # flipping both images left-right flips the gradient field and keeps the loss (symmetric taps)
    sr_f = sr.detach().flip(-1).clone().requires_grad_(True)
    l_f = m(sr_f, hr.flip(-1))
    l_f.backward()
    assert rel_err(l_f.item(), l_all.item()) < 1e-5
    assert (sr_f.grad.flip(-1) - g_all).abs().max().item() < 1e-4 * g_all.abs().max().item()


def test_rejects_what_the_kernels_cannot_do():
    from srgan_st_b200 import StructureTensorLoss
    m = StructureTensorLoss()
    a = torch.rand(1, 3, 16, 16)
    with pytest.raises(RuntimeError):
        m(a, a)                                   # CPU tensors: no fallback
    with pytest.raises(TypeError):
        m(a.cuda().half(), a.cuda().half())
    with pytest.raises(ValueError):
        m(a.cuda()[:, :2], a.cuda()[:, :2])


@pytest.mark.parametrize("sigma,rho", [(1.0, 2.5), (0.3, 1.0), (0.75, 1.5), (0.5, 3.0), (1.0, 1.0)])
def test_other_filter_radii_match_oracle(sigma, rho):
    rng = np.random.default_rng(int(10 * sigma + rho))
    sr = rng.random((2, 3, 100, 152), dtype=np.float32)
    hr = rng.random((2, 3, 100, 152), dtype=np.float32)
    loss, d_sr, d_hr = _run(sr, hr, sigma=sigma, rho=rho)
    ref = O.st_loss(sr, hr, sigma=sigma, rho=rho, want_hr_grad=True)
    assert rel_err(loss, ref["loss"]) < 1e-5
    assert maxnorm_err(d_sr, ref["d_sr"]) < 1e-4 and maxnorm_err(d_hr, ref["d_hr"]) < 1e-4


def test_golden_sigma1_rho25_matches_reference():
    z = golden("st_rand_s1_r25_1x32x40")
    loss, d_sr, d_hr = _run(z["sr"], z["hr"], sigma=1.0, rho=2.5)
    ref = O.st_loss(z["sr"], z["hr"], taps=(z["g"], z["dg"], z["k"]), want_hr_grad=True)
    assert rel_err(loss, z["loss"]) < 1e-5
    assert maxnorm_err(d_sr, ref["d_sr"]) < 1e-4
    assert maxnorm_err(d_sr, z["d_sr"]) < 1e-4 + maxnorm_err(z["d_sr"], ref["d_sr"])


@pytest.mark.parametrize("name,sigma,rho", GENERIC_CASES)
def test_generic_radius_golden_matches_reference(name, sigma, rho):
    """Radii beyond the compiled classes (utils.py:198 is unbounded) run on the generic-radius path: reference
    outputs, fp64 oracle, both gradients."""
    z = golden(name)
    loss, d_sr, d_hr = _run(z["sr"], z["hr"], sigma=sigma, rho=rho)
    ref = O.st_loss(z["sr"], z["hr"], taps=(z["g"], z["dg"], z["k"]), want_hr_grad=True)
    assert rel_err(loss, z["loss"]) < 1e-5 and rel_err(loss, ref["loss"]) < 1e-5
    # Gradients vs the reference's own: 1e-4 widened by the reference's own distance from the float64 truth (the SR-like
    # fixture with sigma 2 / rho 5 is smooth enough that its fp32 gradient wanders; the measured errors are recorded).
    rec = {}
    for tag, ours, orc, refg in (("dsr", d_sr, ref["d_sr"], z["d_sr"]), ("dhr", d_hr, ref["d_hr"], z["d_hr"])):
        e_orc, e_ref, ref_orc = maxnorm_err(ours, orc), maxnorm_err(ours, refg), maxnorm_err(refg, orc)
        rec.update({f"{tag}_vs_fp64": e_orc, f"{tag}_vs_ref": e_ref, f"{tag}_ref_vs_fp64": ref_orc})
        assert e_orc < 1e-4, "grad vs fp64 oracle"
        assert e_ref < 1e-4 + ref_orc, "grad vs reference"
    _record("golden_" + name, loss_vs_ref=rel_err(loss, z["loss"]), loss_vs_fp64=rel_err(loss, ref["loss"]), **rec)


@pytest.mark.parametrize("sigma,rho,shape", [(2.0, 2.0, (2, 100, 152)), (0.5, 4.0, (1, 96, 96)), (3.0, 8.0, (1, 37, 53)),
                                             (1.5, 16.0, (1, 130, 70))])
def test_generic_radius_matches_oracle(sigma, rho, shape):
    """Either radius alone beyond its class, radii larger than the image, the largest radius (rho = 16 -> 64)."""
    rng = np.random.default_rng(int(10 * sigma + rho))
    sr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    hr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    loss, d_sr, d_hr = _run(sr, hr, sigma=sigma, rho=rho)
    ref = O.st_loss(sr, hr, sigma=sigma, rho=rho, want_hr_grad=True)
    assert rel_err(loss, ref["loss"]) < 1e-5
    assert maxnorm_err(d_sr, ref["d_sr"]) < 1e-4 and maxnorm_err(d_hr, ref["d_hr"]) < 1e-4
    # no gradient wanted: nothing saved, same loss; and a second call re-uses the cached workspace
    from srgan_st_b200 import StructureTensorLoss
    with torch.no_grad():
        l2 = StructureTensorLoss(sigma=sigma, rho=rho)(torch.from_numpy(sr).cuda(), torch.from_numpy(hr).cuda())
    assert l2.item() == loss


def test_unsupported_radius_raises():
    from srgan_st_b200 import StructureTensorLoss, StructureTensorPixelLoss, structure_tensor_features
    a = torch.rand(1, 3, 32, 32, device="cuda")
    with pytest.raises(NotImplementedError):
        StructureTensorLoss(sigma=17.0)(a, a)     # radius 68 > 64
    with pytest.raises(NotImplementedError):
        StructureTensorLoss(rho=16.2)(a, a)       # radius 65 > 64
    with pytest.raises(NotImplementedError):
        StructureTensorPixelLoss(rho=4.0)(a, a)   # the fused variant exists for the compiled classes only
    with pytest.raises(NotImplementedError):
        structure_tensor_features(a, sigma=2.0)
