"""CPU oracle for the structure-tensor loss -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference's algorithm (SebastianBitsch/SRGAN-ST,
``loss.py:380-413`` + ``utils.py:194-279``) plus the hand-derived backward pass that the
reference gets from autograd.  It exists to *check* the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product package ``srgan_st_b200`` never does (it fails loudly without its CUDA
library instead).

Pinning: the reference ships no tests or golden vectors for this path.  The
oracle is therefore pinned against outputs of the reference itself, run in the build container
by ``tests/golden/make_golden.py`` and committed as ``tests/golden/st_*.npz``
(``tests/test_oracle_st.py`` replays them).

All arithmetic is float64 by default (``dtype=np.float32`` gives a same-precision port that is
used as the timed CPU baseline).
"""
from __future__ import annotations

import numpy as np

# torchvision.transforms.Grayscale -> _functional_tensor.rgb_to_grayscale (called at loss.py:400-401)
GRAY_R, GRAY_G, GRAY_B = 0.2989, 0.587, 0.114


def gaussian_taps(sigma: float, also_dg: bool = False, radius: int | None = None):
    """utils.py:194-208 ``get_gaussian_kernel`` restated with fp32 numpy.

    radius = max(int(4*sigma + 0.5), 1); phi = exp(-0.5/(sigma^2+1e-12) * x^2) normalised to
    sum 1; dg = phi * -x / (sigma^2 + 1e-12).  Returned as float32 arrays (the reference's
    dtype); they may differ from torch's by 1 ulp of expf, which the golden test bounds.
    """
    if radius is None:
        radius = max(int(4 * sigma + 0.5), 1)
    x = np.arange(-radius, radius + 1, dtype=np.int64)
    sigma2 = (sigma * sigma) + 1e-12
    arg = (np.float32(-0.5 / sigma2) * (x ** 2).astype(np.float32)).astype(np.float32)
    phi = np.exp(arg).astype(np.float32)
    phi = (phi / phi.sum(dtype=np.float32)).astype(np.float32)
    if also_dg:
        dg = ((phi * (-x).astype(np.float32)).astype(np.float32) / np.float32(sigma2)).astype(np.float32)
        return phi, dg
    return phi


def _corr_axis(a: np.ndarray, w: np.ndarray, axis: int) -> np.ndarray:
    """Zero-padded 'same' cross-correlation (not flipped) along one axis: what
    ``F.conv2d(x, w.reshape(...), padding='same')`` does at utils.py:219-230."""
    r = (len(w) - 1) // 2
    pad = [(0, 0)] * a.ndim
    pad[axis] = (r, r)
    ap = np.pad(a, pad)
    out = np.zeros_like(a)
    n = a.shape[axis]
    for i, wi in enumerate(w):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(i, i + n)
        out += wi * ap[tuple(sl)]
    return out


def _corr_axis_T(a: np.ndarray, w: np.ndarray, axis: int) -> np.ndarray:
    """Adjoint of :func:`_corr_axis` = zero-padded correlation with the flipped taps."""
    return _corr_axis(a, w[::-1], axis)


def grayscale(img: np.ndarray) -> np.ndarray:
    """[..., 3, H, W] -> [..., H, W]; torchvision formula (loss.py:400-401)."""
    return (GRAY_R * img[..., 0, :, :] + GRAY_G * img[..., 1, :, :]) + GRAY_B * img[..., 2, :, :]


def structure_tensor(gray: np.ndarray, g, dg, k):
    """utils.py:212-233.  gray [...,H,W] -> (Ix, Iy, Jxx, Jyy, Jxy).

    Ix is the derivative along H (dg applied down the rows, g across), Iy along W.
    """
    ax_h, ax_w = gray.ndim - 2, gray.ndim - 1
    Ix = _corr_axis(_corr_axis(gray, dg, ax_h), g, ax_w)
    Iy = _corr_axis(_corr_axis(gray, g, ax_h), dg, ax_w)
    sm = lambda p: _corr_axis(_corr_axis(p, k, ax_h), k, ax_w)
    return Ix, Iy, sm(Ix * Ix), sm(Iy * Iy), sm(Ix * Iy)


def _clamp_min(x, m):
    # torch.clamp(min=m) propagates NaN; np.maximum does too.
    return np.maximum(x, m)


def st_distance_fields(S1, S2, normalize=True, eps=1e-12):
    """utils.py:236-279 on raw structure tensors S1=(a,b,c) [SR], S2=(e,f,h) [HR].

    Returns a dict with every intermediate the backward pass needs.
    """
    a, b, c = S1
    e, f, h = S2
    with np.errstate(invalid="ignore", divide="ignore"):
        if normalize:  # utils.py:236-239
            q1 = np.sqrt(a * b - c * c + eps)
            q2 = np.sqrt(e * f - h * h + eps)
        else:
            q1 = np.ones_like(a)
            q2 = np.ones_like(e)
        ah, bh, ch = a / q1, b / q1, c / q1
        eh, fh, hh = e / q2, f / q2, h / q2
        # utils.py:248-251
        A = bh * eh - ch * hh
        Bm = ah * fh - ch * hh
        C = bh * hh - ch * fh
        D = ah * hh - ch * eh
        # utils.py:260-265
        T = A + Bm
        disc_raw = T * T - 4.0 * (A * Bm - C * D)
        disc = _clamp_min(disc_raw, eps)
        r = np.sqrt(disc)
        l1_raw = 0.5 * (T - r)
        l2_raw = 0.5 * (T + r)
        # utils.py:275-279
        l1 = _clamp_min(l1_raw, 1.0)
        l2 = _clamp_min(l2_raw, 1.0)
        L1 = np.log(l1)
        L2 = np.log(l2)
        d = np.sqrt(L1 * L1 + L2 * L2 + eps)
    return dict(q1=q1, q2=q2, ah=ah, bh=bh, ch=ch, eh=eh, fh=fh, hh=hh, A=A, Bm=Bm, C=C, D=D,
                T=T, disc_raw=disc_raw, r=r, l1_raw=l1_raw, l2_raw=l2_raw, l1=l1, l2=l2,
                L1=L1, L2=L2, d=d)


def st_distance_backward(S1, S2, F, gd, normalize=True, eps=1e-12, want_hr=False):
    """Adjoint of :func:`st_distance_fields` (what autograd derives for utils.py:236-279).

    gd = dLoss/dd per pixel.  Returns (da, db, dc) w.r.t. the raw SR tensor and, if
    ``want_hr``, (de, df, dh) w.r.t. the raw HR tensor.  Clamp sub-gradients follow torch:
    the gradient passes where input >= min.
    """
    a, b, c = S1
    e, f, h = S2
    with np.errstate(invalid="ignore", divide="ignore"):
        dl1 = gd * F["L1"] / F["d"] / F["l1"] * (F["l1_raw"] >= 1.0)
        dl2 = gd * F["L2"] / F["d"] / F["l2"] * (F["l2_raw"] >= 1.0)
        dT = 0.5 * (dl1 + dl2)
        dr = 0.5 * (dl2 - dl1)
        ddisc = dr / (2.0 * F["r"]) * (F["disc_raw"] >= eps)
        dT = dT + 2.0 * F["T"] * ddisc
        dA = dT - 4.0 * ddisc * F["Bm"]
        dB = dT - 4.0 * ddisc * F["A"]
        dC = 4.0 * ddisc * F["D"]
        dD = 4.0 * ddisc * F["C"]
        ah, bh, ch, eh, fh, hh = (F[n] for n in ("ah", "bh", "ch", "eh", "fh", "hh"))
        # SR side: M = adj(S1^) S2^
        dah = dB * fh + dD * hh
        dbh = dA * eh + dC * hh
        dch = -(dA + dB) * hh - dC * fh - dD * eh

        def through_normalize(x0, x1, x2, g0, g1, g2, q):
            if not normalize:
                return g0, g1, g2
            s = x0 * g0 + x1 * g1 + x2 * g2
            ddet = -s / (2.0 * q ** 3)
            return g0 / q + ddet * x1, g1 / q + ddet * x0, g2 / q - 2.0 * ddet * x2

        out_sr = through_normalize(a, b, c, dah, dbh, dch, F["q1"])
        if not want_hr:
            return out_sr, None
        deh = dA * bh - dD * ch
        dfh = dB * ah - dC * ch
        dhh = -(dA + dB) * ch + dC * bh + dD * ah
        out_hr = through_normalize(e, f, h, deh, dfh, dhh, F["q2"])
        return out_sr, out_hr


def _st_image_backward(Ix, Iy, dS, g, dg, k):
    """dLoss/dgray from dLoss/d(Jxx,Jyy,Jxy): adjoint of utils.py:219-230."""
    ax_h, ax_w = Ix.ndim - 2, Ix.ndim - 1
    smT = lambda p: _corr_axis_T(_corr_axis_T(p, k, ax_w), k, ax_h)
    dPxx, dPyy, dPxy = smT(dS[0]), smT(dS[1]), smT(dS[2])
    dIx = 2.0 * Ix * dPxx + Iy * dPxy
    dIy = 2.0 * Iy * dPyy + Ix * dPxy
    return (_corr_axis_T(_corr_axis_T(dIx, g, ax_w), dg, ax_h)
            + _corr_axis_T(_corr_axis_T(dIy, dg, ax_w), g, ax_h))


def st_loss(sr, hr, sigma=0.5, rho=2.0, normalize=True, taps=None, want_grad=True,
            want_hr_grad=False, dtype=np.float64):
    """StructureTensorLoss.forward (loss.py:399-413) and its backward.

    sr, hr: [B,3,H,W].  ``taps`` = (g, dg, k) overrides :func:`gaussian_taps` (pass the
    reference's own fp32 taps to remove the 1-ulp expf ambiguity).
    Returns dict(loss, d_sr, d_hr, dS_sr, fields...).  The loss is the global mean of the
    per-pixel distance (mean over pixels then over batch == global mean, all images equal size).
    """
    sr = np.asarray(sr, dtype=dtype)
    hr = np.asarray(hr, dtype=dtype)
    if taps is None:
        g, dg = gaussian_taps(sigma, also_dg=True)
        k = gaussian_taps(rho)
    else:
        g, dg, k = taps
    g, dg, k = (np.asarray(t, dtype=dtype) for t in (g, dg, k))
    B, C, H, W = sr.shape
    assert C == 3 and hr.shape == sr.shape
    g1, g2 = grayscale(sr), grayscale(hr)
    Ix1, Iy1, a, b, c = structure_tensor(g1, g, dg, k)
    Ix2, Iy2, e, f, h = structure_tensor(g2, g, dg, k)
    F = st_distance_fields((a, b, c), (e, f, h), normalize)
    P = B * H * W
    out = dict(loss=F["d"].sum(dtype=np.float64) / P, d=F["d"], S_sr=(a, b, c), S_hr=(e, f, h),
               Ix_sr=Ix1, Iy_sr=Iy1)
    if not want_grad:
        return out
    gd = np.full_like(a, 1.0 / P)
    dS1, dS2 = st_distance_backward((a, b, c), (e, f, h), F, gd, normalize, want_hr=want_hr_grad)
    coef = np.array([GRAY_R, GRAY_G, GRAY_B], dtype=dtype).reshape(1, 3, 1, 1)
    out["dS_sr"] = dS1
    out["d_sr"] = coef * _st_image_backward(Ix1, Iy1, dS1, g, dg, k)[:, None]
    if want_hr_grad:
        out["dS_hr"] = dS2
        out["d_hr"] = coef * _st_image_backward(Ix2, Iy2, dS2, g, dg, k)[:, None]
    return out
