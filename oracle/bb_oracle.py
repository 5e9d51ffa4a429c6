"""CPU oracle for the Best-Buddy loss -- TEST INFRASTRUCTURE ONLY.

Two referees:
  * ``bb_forward_c``: ctypes wrapper of oracle/bb_oracle.c, the fp32 restatement with the same
    fixed operation order as the CUDA kernels (indices comparable bit-exactly).
  * ``bb_scores_f64``: the reference's score matrix (loss.py:116-133, utils.py:173-187) in float64
    numpy, used to decide whether a row is a near-tie when comparing with the reference's own
    indices (torch.bmm's summation order is unspecified).
Pinned against tests/golden/bb_*.npz (outputs of the reference itself; tests/test_oracle_bb.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libbb_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "bb_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        _lib = ctypes.CDLL(LIB)
        _lib.bb_oracle_forward.restype = ctypes.c_int
        _lib.bb_oracle_forward_mode.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def pyramid_c(gt: np.ndarray):
    """Bicubic x1/2 and x1/4 levels (loss.py:123,127) with the fixed t=0.5 taps."""
    gt = np.ascontiguousarray(gt, np.float32)
    B, C, H, W = gt.shape
    o2 = np.empty((B, C, H // 2, W // 2), np.float32)
    o4 = np.empty((B, C, H // 4, W // 4), np.float32)
    _load().bb_oracle_pyramid(_fp(gt), B * C, H, W, _fp(o2), _fp(o4))
    return o2, o4


def geometry(H, W):
    N0 = (H // 3) * (W // 3)
    N2 = ((H // 2) // 3) * ((W // 2) // 3)
    N4 = ((H // 4) // 3) * ((W // 4) // 3)
    return N0, N0 + N2 + N4


_MODES = {"patch": 0, "gram": 1, "pst": 2}
_taps_keepalive = None


def pst_taps5(g, dg, k):
    """Central five taps (offsets -2..2) of g, dg, k: all a 3x3 patch image can see."""
    def c5(t):
        t = np.asarray(t, np.float32)
        r = len(t) // 2
        return [t[r + i] if -r <= i <= r else np.float32(0) for i in range(-2, 3)]
    return np.asarray(c5(g) + c5(dg) + c5(k), np.float32)


def _set_pst_taps(taps):
    global _taps_keepalive
    _taps_keepalive = np.ascontiguousarray(pst_taps5(*taps), np.float32)
    _load().bb_oracle_set_pst_taps(_fp(_taps_keepalive))


def describe_c(img, mode="pst", taps=None):
    """fp32 descriptors [B,N,D] of the level-0 patches, same operation order as the CUDA pack kernel."""
    img = np.ascontiguousarray(img, np.float32)
    B, _, H, W = img.shape
    if mode == "pst":
        _set_pst_taps(taps)
    out = np.empty((B, (H // 3) * (W // 3), 9 if mode == "gram" else 27), np.float32)
    _load().bb_oracle_describe(_fp(img), B, H, W, _MODES[mode], _fp(out))
    return out


def bb_forward_c(sr, gt, gt2=None, gt4=None, alpha=1.0, beta=1.0, criterion="l1", mode="patch", taps=None,
                 dist_norm="l2"):
    """mode="patch": BestBuddyLoss (27 raw values); mode="gram": GramLoss (3x3 Gram matrix, loss.py:146-225);
    mode="pst": PatchwiseStructureTensorLoss (loss.py:292-375), taps = (g, dg, k) of utils.get_gaussian_kernel.
    dist_norm: 'l2' (default) or 'l1' (utils.py:166-172), the distance of the search."""
    if mode == "pst":
        _set_pst_taps(taps)
    sr = np.ascontiguousarray(sr, np.float32)
    gt = np.ascontiguousarray(gt, np.float32)
    if gt2 is None:
        gt2, gt4 = pyramid_c(gt)
    gt2 = np.ascontiguousarray(gt2, np.float32)
    gt4 = np.ascontiguousarray(gt4, np.float32)
    B, _, H, W = sr.shape
    N, M = geometry(H, W)
    idx = np.empty((B, N), np.int64)
    best = np.empty((B, N), np.float32)
    second = np.empty((B, N), np.float32)
    loss = ctypes.c_double(0.0)
    rc = _load().bb_oracle_forward_mode(_fp(sr), _fp(gt), _fp(gt2), _fp(gt4), B, H, W, ctypes.c_float(alpha),
                                   ctypes.c_float(beta), (0 if criterion == "l1" else 1) | (0x100 if dist_norm == "l1" else 0),
                                   _MODES[mode],
                                   idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(loss),
                                   _fp(best), _fp(second))
    assert rc == 0
    return dict(idx=idx, loss=loss.value, best=best, second=second)


def unfold3(img):
    """F.unfold(kernel=3, stride=3, pad=0).permute(0,2,1): [B,3,H,W] -> [B,N,27] (loss.py:116-118)."""
    B, C, H, W = img.shape
    ny, nx = H // 3, W // 3
    p = img[:, :, :ny * 3, :nx * 3].reshape(B, C, ny, 3, nx, 3)
    return p.transpose(0, 2, 4, 1, 3, 5).reshape(B, ny * nx, C * 9)


def bb_scores_f64(sr, gt, gt2, gt4, alpha=1.0, beta=1.0, dist_norm="l2"):
    """[B,N,M] float64 scores, the exact-arithmetic version of loss.py:132-133."""
    p1 = unfold3(np.asarray(sr, np.float64))
    p2 = unfold3(np.asarray(gt, np.float64))
    cat = np.concatenate([p2, unfold3(np.asarray(gt2, np.float64)), unfold3(np.asarray(gt4, np.float64))], 1)

    def dist(x, y):
        if dist_norm == "l1":   # utils.py:166-172
            return np.abs(x[:, :, None, :] - y[:, None, :, :]).sum(3)
        d = (x ** 2).sum(2)[:, :, None] + (y ** 2).sum(2)[:, None, :] - 2.0 * np.einsum("bnd,bmd->bnm", x, y)
        return np.clip(d, 0.0, None)

    return alpha * dist(p1, cat) + beta * dist(p2, cat), p1, cat


def bb_backward(sr, cat_sel, criterion="l1"):
    """d loss / d sr for the final criterion only (loss.py:139): sign or 2*diff, /(B*N*27), folded
    back through the non-overlapping unfold.  cat_sel: [B,N,27] selected candidates."""
    sr = np.asarray(sr, np.float64)
    B, C, H, W = sr.shape
    ny, nx = H // 3, W // 3
    diff = unfold3(sr) - cat_sel
    g = (np.sign(diff) if criterion == "l1" else 2.0 * diff) / diff.size
    out = np.zeros_like(sr)
    out[:, :, :ny * 3, :nx * 3] = g.reshape(B, ny, nx, C, 3, 3).transpose(0, 3, 1, 4, 2, 5).reshape(B, C, ny * 3, nx * 3)
    return out


def gram_descriptors(img):
    """[B,3,H,W] -> [B,N,9] Gram matrices of the 3x3x3 patches, float64 (loss.py:180-197)."""
    p = unfold3(np.asarray(img, np.float64))          # [B,N,27] as (c, ky, kx)
    f = p.reshape(p.shape[0], p.shape[1], 3, 9)
    return (np.einsum("bnas,bncs->bnac", f, f) / 27.0).reshape(p.shape[0], p.shape[1], 9)


def gram_backward(sr, sel_desc, criterion="l1"):
    """d GramLoss / d sr given the selected candidate descriptors [B,N,9] (float64)."""
    sr = np.asarray(sr, np.float64)
    B, C, H, W = sr.shape
    ny, nx = H // 3, W // 3
    p = unfold3(sr).reshape(B, ny * nx, 3, 9)
    G1 = np.einsum("bnas,bncs->bnac", p, p) / 27.0
    diff = G1 - sel_desc.reshape(B, ny * nx, 3, 3)
    dG = (np.sign(diff) if criterion == "l1" else 2.0 * diff) / diff.size
    dF = np.einsum("bnac,bncs->bnas", dG + dG.transpose(0, 1, 3, 2), p) / 27.0
    out = np.zeros_like(sr)
    out[:, :, :ny * 3, :nx * 3] = dF.reshape(B, ny, nx, C, 3, 3).transpose(0, 3, 1, 4, 2, 5).reshape(B, C, ny * 3, nx * 3)
    return out


# ---- PatchwiseStructureTensorLoss (loss.py:292-375), float64 -------------------------------------
_GRAY = np.array([0.2989, 0.587, 0.114])


def _band3(w):
    """3x3 matrix M[i][j] = w[j - i + r] of a zero-padded cross-correlation on a length-3 signal."""
    w = np.asarray(w, np.float64)
    r = len(w) // 2
    M = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            if 0 <= j - i + r < len(w):
                M[i, j] = w[j - i + r]
    return M


def _pst_parts(img, taps):
    g, dg, k = taps
    G, DG, K = _band3(g), _band3(dg), _band3(k)
    p = unfold3(np.asarray(img, np.float64))
    p = p.reshape(p.shape[0], p.shape[1], 3, 3, 3)
    gray = np.einsum("c,bncyx->bnyx", _GRAY, p)
    Ix = np.einsum("ij,bnjk,xk->bnix", DG, gray, G)      # (im * dg|) * g-   utils.py:219-220
    Iy = np.einsum("ij,bnjk,xk->bnix", G, gray, DG)      # (im * g|) * dg-   utils.py:221-222
    sm = lambda a: np.einsum("ij,bnjk,xk->bnix", K, a, K)
    J = np.stack([sm(Ix * Ix), sm(Iy * Iy), sm(Ix * Iy)], 2)   # [B,N,3,3,3]
    q = np.sqrt(J[:, :, 0] * J[:, :, 1] - J[:, :, 2] ** 2 + 1e-12)
    return Ix, Iy, J, q, (G, DG, K)


def pst_descriptors(img, taps):
    """[B,3,H,W] -> [B,N,27] float64: normalised structure tensor of every 3x3 patch (loss.py:325-345)."""
    _, _, J, q, _ = _pst_parts(img, taps)
    return (J / q[:, :, None]).reshape(J.shape[0], J.shape[1], 27)


def pst_backward(sr, sel_desc, taps, criterion="l1"):
    """d PatchwiseStructureTensorLoss / d sr given the selected candidate descriptors [B,N,27] (float64)."""
    sr = np.asarray(sr, np.float64)
    B, C, H, W = sr.shape
    ny, nx = H // 3, W // 3
    Ix, Iy, J, q, (G, DG, K) = _pst_parts(sr, taps)
    Dn = J / q[:, :, None]
    diff = Dn - sel_desc.reshape(Dn.shape)
    dD = (np.sign(diff) if criterion == "l1" else 2.0 * diff) / diff.size
    s = (J * dD).sum(2)
    ddet = -s / (2.0 * q ** 3)
    dJ = dD / q[:, :, None]
    dJ[:, :, 0] += ddet * J[:, :, 1]
    dJ[:, :, 1] += ddet * J[:, :, 0]
    dJ[:, :, 2] -= 2.0 * ddet * J[:, :, 2]
    smT = lambda a: np.einsum("ij,bnix,xk->bnjk", K, a, K)
    dP = [smT(dJ[:, :, c]) for c in range(3)]
    dIx = 2 * Ix * dP[0] + Iy * dP[2]
    dIy = 2 * Iy * dP[1] + Ix * dP[2]
    dgray = np.einsum("ij,bnix,xk->bnjk", DG, dIx, G) + np.einsum("ij,bnix,xk->bnjk", G, dIy, DG)
    dF = _GRAY[None, None, :, None, None] * dgray[:, :, None]          # [B,N,3,3,3]
    out = np.zeros_like(sr)
    out[:, :, :ny * 3, :nx * 3] = dF.reshape(B, ny, nx, C, 3, 3).transpose(0, 3, 1, 4, 2, 5).reshape(B, C, ny * 3, nx * 3)
    return out


# ---- gradient w.r.t. gt (loss.py:136-139: the gather of the selected candidates is differentiable in p2_cat) ----
def _fold3(p, H, W):
    """inverse of unfold3 for non-overlapping 3x3 patches: [B,N,27] -> [B,3,H,W] (pixels outside every patch = 0)."""
    B = p.shape[0]
    ny, nx = H // 3, W // 3
    out = np.zeros((B, 3, H, W), p.dtype)
    out[:, :, :ny * 3, :nx * 3] = p.reshape(B, ny, nx, 3, 3, 3).transpose(0, 3, 1, 4, 2, 5).reshape(B, 3, ny * 3, nx * 3)
    return out


def _bicubic_adjoint(d, H, W, scale):
    """adjoint of F.interpolate(gt, scale_factor=1/scale, mode='bicubic', align_corners=False) for scale 2 or 4:
    taps (-0.09375, 0.59375, 0.59375, -0.09375) on source rows / columns 2i-1..2i+2 (x1/2) or 4i..4i+3 (x1/4), indices
    clamped to the image (bb_oracle.c pyramid, F.interpolate)."""
    w = np.array([-0.09375, 0.59375, 0.59375, -0.09375])
    B, C, Ho, Wo = d.shape
    out = np.zeros((B, C, H, W), np.float64)
    ys = np.arange(Ho)[:, None] * scale + (0 if scale == 4 else -1) + np.arange(4)[None, :]
    xs = np.arange(Wo)[:, None] * scale + (0 if scale == 4 else -1) + np.arange(4)[None, :]
    ys, xs = np.clip(ys, 0, H - 1), np.clip(xs, 0, W - 1)
    for i in range(4):
        for k in range(4):
            np.add.at(out, (slice(None), slice(None), ys[:, i][:, None], xs[:, k][None, :]), d * (w[i] * w[k]))
    return out


def patch_backward_gt(sr, gt, gt2, gt4, idx, mode="patch", taps=None, criterion="l1"):
    """d loss / d gt of BestBuddyLoss / GramLoss / PatchwiseStructureTensorLoss given the argmin indices [B,N]
    (float64).  The final criterion is symmetric, so the gradient w.r.t. a selected candidate is the SR-side formula
    (bb_backward / gram_backward / pst_backward) with the two operands swapped; it is accumulated over the queries
    that selected the candidate and folded back through the pyramid."""
    sr = np.asarray(sr, np.float64)
    gt = np.asarray(gt, np.float64)
    gt2 = np.asarray(gt2, np.float64)
    gt4 = np.asarray(gt4, np.float64)
    B, _, H, W = sr.shape
    ny, nx = H // 3, W // 3
    N = ny * nx
    cat = np.concatenate([unfold3(gt), unfold3(gt2), unfold3(gt4)], 1)             # [B,M,27] candidate patches
    sel = np.take_along_axis(cat, idx[:, :, None].astype(np.int64), 1)            # [B,N,27]
    sel_img = _fold3(sel, 3 * ny, 3 * nx)                                          # the selected candidates as an image
    sr_c = sr[:, :, :3 * ny, :3 * nx]
    if mode == "patch":
        g_img = bb_backward(sel_img, unfold3(sr_c), criterion)
    elif mode == "gram":
        g_img = gram_backward(sel_img, gram_descriptors(sr_c), criterion)
    else:
        g_img = pst_backward(sel_img, pst_descriptors(sr_c, taps), taps, criterion)
    g_sel = unfold3(g_img)                                                         # [B,N,27] per-query candidate gradients
    g_cat = np.zeros_like(cat)
    for b in range(B):
        np.add.at(g_cat[b], idx[b], g_sel[b])
    N0, N2 = N, (gt2.shape[2] // 3) * (gt2.shape[3] // 3)
    d0 = _fold3(g_cat[:, :N0], H, W)
    d2 = _fold3(g_cat[:, N0:N0 + N2], gt2.shape[2], gt2.shape[3])
    d4 = _fold3(g_cat[:, N0 + N2:], gt4.shape[2], gt4.shape[3])
    return d0 + _bicubic_adjoint(d2, H, W, 2) + _bicubic_adjoint(d4, H, W, 4)
