"""Generate golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference hard-codes ``.cuda()`` in ``utils.py:206,208``; on a CPU-only box the single shim
``torch.Tensor.cuda = identity`` lets ``loss.py`` run unchanged.  Outputs
are small ``.npz`` files committed next to this script; ``tests/`` replays them against the
oracle (CPU) and against the CUDA path (GPU).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SRST_REFERENCE", "/root/reference")


def _import_reference():
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # the one shim
    sys.path.insert(0, REF)
    import loss as ref_loss  # noqa: E402
    import utils as ref_utils  # noqa: E402
    return ref_loss, ref_utils


def _inputs(kind, B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "rand":
        sr = torch.rand(B, 3, H, W, generator=g)
        hr = torch.rand(B, 3, H, W, generator=g)
    elif kind == "srlike":
        # HR = k/255 low-passed noise, SR = blurred HR + noise, saturated to [0,1] (model.py:150)
        hr = torch.randint(0, 256, (B, 3, H, W), generator=g).float()
        hr = torch.nn.functional.avg_pool2d(hr, 3, 1, 1, count_include_pad=False).round() / 255
        sr = torch.nn.functional.avg_pool2d(hr, 3, 1, 1, count_include_pad=False)
        sr = (sr + 0.05 * torch.randn(B, 3, H, W, generator=g) + 0.1).clamp(0, 1)
    elif kind == "same":
        hr = torch.rand(B, 3, H, W, generator=g)
        sr = hr.clone()
    else:
        raise ValueError(kind)
    return sr.contiguous(), hr.contiguous()


ST_CASES = [
    # name, kind, B, H, W, seed, sigma, rho, normalize
    ("st_rand_2x24x36", "rand", 2, 24, 36, 1, 0.5, 2.0, True),
    ("st_rand_1x96x96", "rand", 1, 96, 96, 2, 0.5, 2.0, True),
    ("st_srlike_2x40x52", "srlike", 2, 40, 52, 3, 0.5, 2.0, True),
    ("st_same_1x24x24", "same", 1, 24, 24, 4, 0.5, 2.0, True),
    ("st_rand_s1_r25_1x32x40", "rand", 1, 32, 40, 5, 1.0, 2.5, True),
    ("st_rand_nonorm_1x24x36", "rand", 1, 24, 36, 6, 0.5, 2.0, False),
    ("st_rand_ragged_1x37x53", "rand", 1, 37, 53, 7, 0.5, 2.0, True),
    # radii beyond the compiled classes (r_sigma 6 / 8, r_rho 16 / 20): the generic-radius path
    ("st_rand_s15_r4_1x40x52", "rand", 1, 40, 52, 8, 1.5, 4.0, True),
    ("st_srlike_s2_r5_2x33x45", "srlike", 2, 33, 45, 9, 2.0, 5.0, True),
]

BB_CASES = [
    # name, kind, B, H, W, seed, alpha, beta, criterion
    ("bb_rand_2x24x24", "rand", 2, 24, 24, 11, 1.0, 1.0, "l1"),
    ("bb_rand_1x48x36", "rand", 1, 48, 36, 12, 1.0, 1.0, "l1"),
    ("bb_srlike_2x48x48", "srlike", 2, 48, 48, 13, 1.0, 1.0, "l1"),
    ("bb_rand_ab_1x36x36", "rand", 1, 36, 36, 14, 0.5, 2.0, "l2"),
    ("bb_same_1x24x24", "same", 1, 24, 24, 15, 1.0, 1.0, "l1"),
]


def make_st(ref_loss, ref_utils):
    for name, kind, B, H, W, seed, sigma, rho, norm in ST_CASES:
        sr, hr = _inputs(kind, B, H, W, seed)
        sr.requires_grad_(True)
        hr.requires_grad_(True)
        m = ref_loss.StructureTensorLoss(sigma=sigma, rho=rho, normalize=norm)
        loss = m(sr, hr)
        loss.backward()
        g, dg = ref_utils.get_gaussian_kernel(sigma, also_dg=True)
        k = ref_utils.get_gaussian_kernel(rho)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            sr=sr.detach().numpy(), hr=hr.detach().numpy(), loss=np.float32(loss.item()),
            d_sr=sr.grad.numpy(), d_hr=hr.grad.numpy(), g=g.numpy(), dg=dg.numpy(), k=k.numpy(),
            sigma=np.float64(sigma), rho=np.float64(rho), normalize=np.bool_(norm))
        print(f"{name}: loss={loss.item():.8g} |d_sr|max={sr.grad.abs().max().item():.4g}")


def make_bb(ref_loss, ref_utils):
    import torch.nn.functional as F
    for name, kind, B, H, W, seed, alpha, beta, crit in BB_CASES:
        sr, hr = _inputs(kind, B, H, W, seed)
        sr.requires_grad_(True)
        m = ref_loss.BestBuddyLoss(alpha=alpha, beta=beta, criterion=crit)
        loss = m(sr, hr)
        loss.backward()
        # Re-derive the indices and the top-2 score gap with the reference's own helpers
        # (loss.py:116-135) so the tests can apply the near-tie protocol.
        with torch.no_grad():
            unf = lambda t: F.unfold(t, kernel_size=3, padding=0, stride=3).permute(0, 2, 1).contiguous()
            p1, p2 = unf(sr), unf(hr)
            hr2 = F.interpolate(hr, scale_factor=0.5, mode="bicubic", align_corners=False)
            hr4 = F.interpolate(hr, scale_factor=0.25, mode="bicubic", align_corners=False)
            cat = torch.cat([p2, unf(hr2), unf(hr4)], 1)
            score = alpha * ref_utils.batch_pairwise_distance(p1, cat, "l2") \
                + beta * ref_utils.batch_pairwise_distance(p2, cat, "l2")
            w, ind = torch.min(score, dim=2)
            top2 = torch.topk(score, 2, dim=2, largest=False).values
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            sr=sr.detach().numpy(), hr=hr.detach().numpy(), loss=np.float32(loss.item()),
            d_sr=sr.grad.numpy(), ind=ind.numpy(), top2=top2.numpy(),
            hr2=hr2.numpy(), hr4=hr4.numpy(),
            alpha=np.float64(alpha), beta=np.float64(beta), criterion=np.str_(crit))
        print(f"{name}: loss={loss.item():.8g} N={p1.shape[1]} M={cat.shape[1]}")


GRAM_CASES = [
    ("gram_rand_2x24x24", "rand", 2, 24, 24, 21, 1.0, 1.0, "l1"),
    ("gram_srlike_2x48x36", "srlike", 2, 48, 36, 22, 1.0, 1.0, "l1"),
    ("gram_rand_ab_1x36x36", "rand", 1, 36, 36, 23, 0.5, 2.0, "l2"),
]


def make_gram(ref_loss, ref_utils):
    import torch.nn.functional as F
    for name, kind, B, H, W, seed, alpha, beta, crit in GRAM_CASES:
        sr, hr = _inputs(kind, B, H, W, seed)
        sr.requires_grad_(True)
        m = ref_loss.GramLoss(alpha=alpha, beta=beta, criterion=crit)
        loss = m(sr, hr)
        loss.backward()
        with torch.no_grad():  # indices / top-2 gaps with the reference's own helpers (loss.py:200-218)
            p1, p2 = m.compute_patches(sr), m.compute_patches(hr)
            hr2 = F.interpolate(hr, scale_factor=0.5, mode="bicubic", align_corners=False)
            hr4 = F.interpolate(hr, scale_factor=0.25, mode="bicubic", align_corners=False)
            cat = torch.cat([p2, m.compute_patches(hr2), m.compute_patches(hr4)], 1)
            score = alpha * ref_utils.batch_pairwise_distance(p1, cat, "l2") \
                + beta * ref_utils.batch_pairwise_distance(p2, cat, "l2")
            _, ind = torch.min(score, dim=2)
            top2 = torch.topk(score, 2, dim=2, largest=False).values
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            sr=sr.detach().numpy(), hr=hr.detach().numpy(), loss=np.float32(loss.item()),
            d_sr=sr.grad.numpy(), ind=ind.numpy(), top2=top2.numpy(), hr2=hr2.numpy(), hr4=hr4.numpy(),
            p1=p1.numpy(), alpha=np.float64(alpha), beta=np.float64(beta), criterion=np.str_(crit))
        print(f"{name}: loss={loss.item():.8g} N={p1.shape[1]} M={cat.shape[1]}")


PST_CASES = [
    # name, kind, B, H, W, seed, sigma, rho, alpha, beta, criterion
    ("pst_rand_2x24x24", "rand", 2, 24, 24, 31, 0.5, 2.0, 1.0, 1.0, "l1"),
    ("pst_srlike_2x48x36", "srlike", 2, 48, 36, 32, 0.5, 2.0, 1.0, 1.0, "l1"),
    ("pst_rand_ab_s1_1x36x36", "rand", 1, 36, 36, 33, 1.0, 2.5, 0.5, 2.0, "l2"),
]


def make_pst(ref_loss, ref_utils):
    import torch.nn.functional as F
    for name, kind, B, H, W, seed, sigma, rho, alpha, beta, crit in PST_CASES:
        sr, hr = _inputs(kind, B, H, W, seed)
        sr.requires_grad_(True)
        m = ref_loss.PatchwiseStructureTensorLoss(sigma=sigma, rho=rho, alpha=alpha, beta=beta, criterion=crit)
        loss = m(sr, hr)
        loss.backward()
        with torch.no_grad():  # indices / top-2 gaps with the reference's own helpers (loss.py:347-369)
            p1, p2 = m.compute_patches(sr), m.compute_patches(hr)
            hr2 = F.interpolate(hr, scale_factor=0.5, mode="bicubic", align_corners=False)
            hr4 = F.interpolate(hr, scale_factor=0.25, mode="bicubic", align_corners=False)
            cat = torch.cat([p2, m.compute_patches(hr2), m.compute_patches(hr4)], 1)
            score = alpha * ref_utils.batch_pairwise_distance(p1, cat, "l2") \
                + beta * ref_utils.batch_pairwise_distance(p2, cat, "l2")
            _, ind = torch.min(score, dim=2)
            top2 = torch.topk(score, 2, dim=2, largest=False).values
        g, dg = ref_utils.get_gaussian_kernel(sigma, also_dg=True)
        k = ref_utils.get_gaussian_kernel(rho)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            sr=sr.detach().numpy(), hr=hr.detach().numpy(), loss=np.float32(loss.item()),
            d_sr=sr.grad.numpy(), ind=ind.numpy(), top2=top2.numpy(), hr2=hr2.numpy(), hr4=hr4.numpy(),
            p1=p1.numpy(), g=g.numpy(), dg=dg.numpy(), k=k.numpy(), sigma=np.float64(sigma), rho=np.float64(rho),
            alpha=np.float64(alpha), beta=np.float64(beta), criterion=np.str_(crit))
        print(f"{name}: loss={loss.item():.8g} N={p1.shape[1]} M={cat.shape[1]}")


GT_GRAD_CASES = [
    # gradient w.r.t. gt (loss.py:136-139): same inputs as the fixture named first; only d_gt (and the loss) is stored
    ("bb_rand_2x24x24", "bb", "rand", 2, 24, 24, 11, dict(alpha=1.0, beta=1.0, criterion="l1")),
    ("bb_srlike_2x48x48", "bb", "srlike", 2, 48, 48, 13, dict(alpha=1.0, beta=1.0, criterion="l1")),
    ("bb_rand_ab_1x36x36", "bb", "rand", 1, 36, 36, 14, dict(alpha=0.5, beta=2.0, criterion="l2")),
    ("gram_srlike_2x48x36", "gram", "srlike", 2, 48, 36, 22, dict(alpha=1.0, beta=1.0, criterion="l1")),
    ("gram_rand_ab_1x36x36", "gram", "rand", 1, 36, 36, 23, dict(alpha=0.5, beta=2.0, criterion="l2")),
    ("pst_rand_2x24x24", "pst", "rand", 2, 24, 24, 31, dict(sigma=0.5, rho=2.0, alpha=1.0, beta=1.0, criterion="l1")),
    ("pst_rand_ab_s1_1x36x36", "pst", "rand", 1, 36, 36, 33, dict(sigma=1.0, rho=2.5, alpha=0.5, beta=2.0, criterion="l2")),
]


def make_gt_grad(ref_loss, ref_utils):
    cls = {"bb": ref_loss.BestBuddyLoss, "gram": ref_loss.GramLoss, "pst": ref_loss.PatchwiseStructureTensorLoss}
    for name, which, kind, B, H, W, seed, kw in GT_GRAD_CASES:
        sr, hr = _inputs(kind, B, H, W, seed)
        sr.requires_grad_(True)
        hr.requires_grad_(True)
        loss = cls[which](**kw)(sr, hr)
        loss.backward()
        np.savez_compressed(os.path.join(HERE, name + "_dgt.npz"), loss=np.float32(loss.item()), d_gt=hr.grad.numpy(),
                            d_sr=sr.grad.numpy())
        print(f"{name}_dgt: loss={loss.item():.8g} |d_gt|max={hr.grad.abs().max().item():.4g}")


L1_CASES = [
    # dist_norm='l1' (utils.py:166-172) for the three patch losses: name, module, kind, B, H, W, seed, alpha, beta, criterion
    ("bbl1_rand_2x24x24", "bb", "rand", 2, 24, 24, 41, 1.0, 1.0, "l1"),
    ("bbl1_srlike_2x48x36", "bb", "srlike", 2, 48, 36, 42, 0.5, 2.0, "l2"),
    ("graml1_rand_1x36x36", "gram", "rand", 1, 36, 36, 43, 1.0, 1.0, "l1"),
    ("pstl1_rand_2x24x24", "pst", "rand", 2, 24, 24, 44, 1.0, 1.0, "l1"),
]


def make_l1(ref_loss, ref_utils):
    import torch.nn.functional as F
    for name, which, kind, B, H, W, seed, alpha, beta, crit in L1_CASES:
        sr, hr = _inputs(kind, B, H, W, seed)
        sr.requires_grad_(True)
        if which == "bb":
            m = ref_loss.BestBuddyLoss(alpha=alpha, beta=beta, dist_norm="l1", criterion=crit)
            desc = lambda t: F.unfold(t, kernel_size=3, padding=0, stride=3).permute(0, 2, 1).contiguous()
        elif which == "gram":
            m = ref_loss.GramLoss(alpha=alpha, beta=beta, dist_norm="l1", criterion=crit)
            desc = m.compute_patches
        else:
            m = ref_loss.PatchwiseStructureTensorLoss(alpha=alpha, beta=beta, dist_norm="l1", criterion=crit)
            desc = m.compute_patches
        loss = m(sr, hr)
        loss.backward()
        with torch.no_grad():
            p1, p2 = desc(sr), desc(hr)
            hr2 = F.interpolate(hr, scale_factor=0.5, mode="bicubic", align_corners=False)
            hr4 = F.interpolate(hr, scale_factor=0.25, mode="bicubic", align_corners=False)
            cat = torch.cat([p2, desc(hr2), desc(hr4)], 1)
            score = alpha * ref_utils.batch_pairwise_distance(p1, cat, "l1") \
                + beta * ref_utils.batch_pairwise_distance(p2, cat, "l1")
            _, ind = torch.min(score, dim=2)
            top2 = torch.topk(score, 2, dim=2, largest=False).values
        extra = {}
        if which == "pst":
            g, dg = ref_utils.get_gaussian_kernel(0.5, also_dg=True)
            extra = dict(g=g.numpy(), dg=dg.numpy(), k=ref_utils.get_gaussian_kernel(2.0).numpy())
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            sr=sr.detach().numpy(), hr=hr.detach().numpy(), loss=np.float32(loss.item()),
            d_sr=sr.grad.numpy(), ind=ind.numpy(), top2=top2.numpy(), hr2=hr2.numpy(), hr4=hr4.numpy(),
            alpha=np.float64(alpha), beta=np.float64(beta), criterion=np.str_(crit), which=np.str_(which), **extra)
        print(f"{name}: loss={loss.item():.8g} N={p1.shape[1]} M={cat.shape[1]}")


GEOM_CASES = [
    # BestBuddyLoss with a non-default patch geometry (loss.py:116-129 use self.ksize / self.pad / self.stride):
    # name, kind, B, H, W, seed, ksize, pad, stride, alpha, beta, dist_norm, criterion
    ("bbg_k4p1s2_rand_2x24x28", "rand", 2, 24, 28, 51, 4, 1, 2, 1.0, 1.0, "l2", "l1"),     # overlapping, zero-padded
    ("bbg_k3p1s1_srlike_1x20x24", "srlike", 1, 20, 24, 52, 3, 1, 1, 0.5, 2.0, "l2", "l2"),  # every pixel in nine patches
    ("bbg_k5p0s5_rand_2x40x45", "rand", 2, 40, 45, 53, 5, 0, 5, 1.0, 1.0, "l2", "l1"),     # 75-dim patches, ragged levels
    ("bbg_k2p0s3_rand_1x24x24", "rand", 1, 24, 24, 54, 2, 0, 3, 1.0, 1.0, "l1", "l1"),     # gaps between patches, l1 search
]


def make_geom(ref_loss, ref_utils):
    import torch.nn.functional as F
    for name, kind, B, H, W, seed, ks, pad, st, alpha, beta, dn, crit in GEOM_CASES:
        sr, hr = _inputs(kind, B, H, W, seed)
        sr.requires_grad_(True)
        hr.requires_grad_(True)   # the gather of the selected candidates is differentiable in gt (loss.py:136-139)
        m = ref_loss.BestBuddyLoss(alpha=alpha, beta=beta, ksize=ks, pad=pad, stride=st, dist_norm=dn, criterion=crit)
        loss = m(sr, hr)
        loss.backward()
        with torch.no_grad():
            unf = lambda t: F.unfold(t, kernel_size=ks, padding=pad, stride=st).permute(0, 2, 1).contiguous()
            p1, p2 = unf(sr), unf(hr)
            hr2 = F.interpolate(hr, scale_factor=0.5, mode="bicubic", align_corners=False)
            hr4 = F.interpolate(hr, scale_factor=0.25, mode="bicubic", align_corners=False)
            cat = torch.cat([p2, unf(hr2), unf(hr4)], 1)
            score = alpha * ref_utils.batch_pairwise_distance(p1, cat, dn) \
                + beta * ref_utils.batch_pairwise_distance(p2, cat, dn)
            _, ind = torch.min(score, dim=2)
            top2 = torch.topk(score, 2, dim=2, largest=False).values
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            sr=sr.detach().numpy(), hr=hr.detach().numpy(), loss=np.float32(loss.item()),
            d_sr=sr.grad.numpy(), d_gt=hr.grad.numpy(), ind=ind.numpy(), top2=top2.numpy(), hr2=hr2.numpy(), hr4=hr4.numpy(),
            alpha=np.float64(alpha), beta=np.float64(beta), criterion=np.str_(crit), dist_norm=np.str_(dn),
            ksize=np.int64(ks), pad=np.int64(pad), stride=np.int64(st))
        print(f"{name}: loss={loss.item():.8g} N={p1.shape[1]} M={cat.shape[1]} D={p1.shape[2]}")


if __name__ == "__main__":
    torch.set_num_threads(1)  # fixed summation order inside MKL for reproducible fixtures
    rl, ru = _import_reference()
    which = sys.argv[1:] or ["st", "bb", "gram", "pst", "l1", "gtgrad", "geom"]
    if "st" in which:
        make_st(rl, ru)
    if "bb" in which:
        make_bb(rl, ru)
    if "gram" in which:
        make_gram(rl, ru)
    if "pst" in which:
        make_pst(rl, ru)
    if "l1" in which:
        make_l1(rl, ru)
    if "gtgrad" in which:
        make_gt_grad(rl, ru)
    if "geom" in which:
        make_geom(rl, ru)
