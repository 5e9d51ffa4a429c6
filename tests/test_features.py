"""Structure-tensor features (srst_st_features: smoothed tensor, closed-form eigenvalues, orientation, coherence --
the diagnostic output BASELINE.json's north_star names).  The reference computes them only in a notebook through a
third-party package that is not in its checkout (parity unpinned), so the check is against first
principles: J against the oracle's structure tensor (which IS pinned on the reference, utils.py:212-233), the
eigen-decomposition against numpy.linalg.eigh in float64."""
import ctypes

import numpy as np
import pytest

from oracle import st_oracle as O
from tests.helpers import emu_lib


def _expected(img, sigma=0.5, rho=2.0):
    g, dg = O.gaussian_taps(sigma, True)
    k = O.gaussian_taps(rho)
    gray = O.grayscale(np.asarray(img, np.float64))
    _, _, a, b, c = O.structure_tensor(gray, g.astype(np.float64), dg.astype(np.float64), k.astype(np.float64))
    J = np.stack([a, b, c], 1)                                            # [B,3,H,W]: Jxx, Jyy, Jxy
    M = np.stack([np.stack([a, c], -1), np.stack([c, b], -1)], -2)        # [[Jxx, Jxy], [Jxy, Jyy]]
    w, v = np.linalg.eigh(M)                                               # ascending eigenvalues
    vmax = v[..., :, 1]                                                    # eigenvector of the large eigenvalue (H, W comps)
    return J, w, vmax, (g, dg, k)


def _check(out, img):
    J, w, vmax, _ = _expected(img)
    assert np.abs(out["J"] - J).max() <= 2e-6 * np.abs(J).max()
    scale = np.abs(w).max()
    assert np.abs(out["eig"][:, 0] - w[..., 0]).max() <= 3e-6 * scale
    assert np.abs(out["eig"][:, 1] - w[..., 1]).max() <= 3e-6 * scale
    aniso = w[..., 1] - w[..., 0] > 1e-3 * scale                           # orientation is defined where the tensor is anisotropic
    ang_ref = np.arctan2(vmax[..., 1], vmax[..., 0])                       # angle against the H axis (component 0)
    d = np.abs(np.angle(np.exp(2j * (out["orient"] - ang_ref))))[aniso] / 2  # eigenvectors are defined up to sign: compare mod pi
    assert d.max() < 2e-3
    coh_ref = np.where(w[..., 1] > 0, 1 - w[..., 0] / np.where(w[..., 1] > 0, w[..., 1], 1), 0)
    big = w[..., 1] > 1e-3 * scale
    assert np.abs(out["coher"] - coh_ref)[big].max() < 2e-3
    assert (out["coher"] >= -1e-6).all() and (out["coher"] <= 1 + 1e-6).all()
    assert (np.abs(out["orient"]) <= np.pi / 2 + 1e-6).all()


@pytest.mark.parametrize("shape", [(2, 3, 40, 72), (1, 3, 37, 53)])
def test_emulated_features_match_first_principles(shape):
    lib = emu_lib()
    rng = np.random.default_rng(shape[2])
    img = rng.random(shape, dtype=np.float32)
    B, _, H, W = shape
    g, dg = O.gaussian_taps(0.5, True)
    k = O.gaussian_taps(2.0)
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    out = dict(J=np.full((B, 3, H, W), np.nan, np.float32), eig=np.full((B, 2, H, W), np.nan, np.float32),
               orient=np.full((B, H, W), np.nan, np.float32), coher=np.full((B, H, W), np.nan, np.float32))
    rc = lib.srst_st_features(vp(img), B, H, W, fp(g), fp(dg), len(g) // 2, fp(k), len(k) // 2, vp(out["J"]), vp(out["eig"]),
                              vp(out["orient"]), vp(out["coher"]), None)
    assert rc == 0
    assert not any(np.isnan(v).any() for v in out.values())
    _check(out, img)
    assert lib.srst_st_features(vp(img), B, H, W, fp(g), fp(dg), len(g) // 2, fp(k), len(k) // 2, None, None, None, None,
                                None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 3, 96, 96), (1, 3, 333, 517)])
def test_gpu_features(shape):
    import torch
    from srgan_st_b200 import structure_tensor_features
    rng = np.random.default_rng(shape[3])
    img = rng.random(shape, dtype=np.float32)
    f = structure_tensor_features(torch.from_numpy(img).cuda())
    torch.cuda.synchronize()
    out = dict(J=f["J"].cpu().numpy(), eig=f["eigenvalues"].cpu().numpy(), orient=f["orientation"].cpu().numpy(),
               coher=f["coherence"].cpu().numpy())
    _check(out, img)
    # an image that only varies along W: the gradient eigenvector lies along W, i.e. at +-pi/2 against the H axis
    ramp = torch.linspace(0, 1, 96).view(1, 1, 1, 96).expand(1, 3, 96, 96).contiguous().cuda()
    f = structure_tensor_features(ramp)
    mid = f["orientation"][0, 20:76, 20:76].abs()
    assert (mid - np.pi / 2).abs().max().item() < 1e-3 and f["coherence"][0, 20:76, 20:76].min().item() > 0.999
    with pytest.raises(TypeError):
        structure_tensor_features(torch.rand(1, 3, 8, 8))
