"""CPU oracle for BestBuddyLoss with a non-default patch geometry -- TEST INFRASTRUCTURE ONLY.

``bbg_forward_c`` wraps oracle/bbg_oracle.c (fp32, the CUDA path's fixed operation order: indices bit-comparable);
``bbg_scores_f64`` is the reference's score matrix (loss.py:116-133) in float64 numpy for the near-tie protocol.
Pinned against tests/golden/bbg_*.npz, which are outputs of the reference itself (tests/golden/make_golden.py geom).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import bb_oracle as _bb

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libbbg_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "bbg_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        _lib = ctypes.CDLL(LIB)
        _lib.bbg_oracle_forward.restype = ctypes.c_int
    return _lib


def npatch(size, k, p, s):
    span = size + 2 * p - k
    return 0 if span < 0 else span // s + 1


def geometry(H, W, k, p, s):
    n = [npatch(h, k, p, s) * npatch(w, k, p, s) for h, w in ((H, W), (H // 2, W // 2), (H // 4, W // 4))]
    return n[0], sum(n)


def bbg_forward_c(sr, gt, gt2=None, gt4=None, ksize=3, pad=0, stride=3, alpha=1.0, beta=1.0, criterion="l1",
                  dist_norm="l2"):
    sr = np.ascontiguousarray(sr, np.float32)
    gt = np.ascontiguousarray(gt, np.float32)
    if gt2 is None:
        gt2, gt4 = _bb.pyramid_c(gt)
    gt2 = np.ascontiguousarray(gt2, np.float32)
    gt4 = np.ascontiguousarray(gt4, np.float32)
    B, _, H, W = sr.shape
    N, _ = geometry(H, W, ksize, pad, stride)
    idx = np.empty((B, N), np.int64)
    best = np.empty((B, N), np.float32)
    second = np.empty((B, N), np.float32)
    d_sr = np.empty_like(sr)
    loss = ctypes.c_double(0.0)
    fp = _bb._fp
    rc = _load().bbg_oracle_forward(fp(sr), fp(gt), fp(gt2), fp(gt4), B, H, W, int(ksize), int(pad), int(stride),
                                    ctypes.c_float(alpha), ctypes.c_float(beta),
                                    (0 if criterion == "l1" else 1) | (0x100 if dist_norm == "l1" else 0),
                                    idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(loss), fp(best),
                                    fp(second), fp(d_sr))
    assert rc == 0, rc
    return dict(idx=idx, loss=loss.value, best=best, second=second, d_sr=d_sr)


def unfold(img, k, p, s):
    """F.unfold(kernel_size=k, padding=p, stride=s).permute(0,2,1): [B,3,H,W] -> [B,N,3*k*k]."""
    B, C, H, W = img.shape
    ny, nx = npatch(H, k, p, s), npatch(W, k, p, s)
    pad = np.zeros((B, C, H + 2 * p, W + 2 * p), img.dtype)
    pad[:, :, p:p + H, p:p + W] = img
    out = np.empty((B, ny * nx, C, k, k), img.dtype)
    for py in range(ny):
        for px in range(nx):
            out[:, py * nx + px] = pad[:, :, py * s:py * s + k, px * s:px * s + k]
    return out.reshape(B, ny * nx, C * k * k)


def bbg_scores_f64(sr, gt, gt2, gt4, ksize, pad, stride, alpha=1.0, beta=1.0, dist_norm="l2"):
    u = lambda t: unfold(np.asarray(t, np.float64), ksize, pad, stride)
    p1, p2 = u(sr), u(gt)
    cat = np.concatenate([p2, u(gt2), u(gt4)], 1)

    def dist(x, y):
        if dist_norm == "l1":
            return np.abs(x[:, :, None, :] - y[:, None, :, :]).sum(3)
        d = (x ** 2).sum(2)[:, :, None] + (y ** 2).sum(2)[:, None, :] - 2.0 * np.einsum("bnd,bmd->bnm", x, y)
        return np.clip(d, 0.0, None)

    return alpha * dist(p1, cat) + beta * dist(p2, cat), p1, cat
