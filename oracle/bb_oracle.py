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


def bb_forward_c(sr, gt, gt2=None, gt4=None, alpha=1.0, beta=1.0, criterion="l1", mode="patch"):
    """mode="patch": BestBuddyLoss (27 raw values); mode="gram": GramLoss (3x3 Gram matrix, loss.py:146-225)."""
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
                                   ctypes.c_float(beta), 0 if criterion == "l1" else 1, 0 if mode == "patch" else 1,
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


def bb_scores_f64(sr, gt, gt2, gt4, alpha=1.0, beta=1.0):
    """[B,N,M] float64 scores, the exact-arithmetic version of loss.py:132-133."""
    p1 = unfold3(np.asarray(sr, np.float64))
    p2 = unfold3(np.asarray(gt, np.float64))
    cat = np.concatenate([p2, unfold3(np.asarray(gt2, np.float64)), unfold3(np.asarray(gt4, np.float64))], 1)

    def dist(x, y):
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
