"""CPU: execute the real ST kernel sources under the test-only host emulation (tests/emu) and
compare with the oracle.  This checks tiling/halo/index logic of every compiled tile shape without
a GPU; numerical parity of the sm_100a build itself is the job of tests/test_st_gpu.py."""
import os

import numpy as np
import pytest

from oracle import st_oracle as O
from tests.helpers import emu_lib, emu_st, golden, maxnorm_err, rel_err


@pytest.fixture(scope="module")
def lib():
    return emu_lib()


@pytest.fixture(params=[0, 1, 2, 3, 4, 10])
def tile_cfg(request, monkeypatch):
    monkeypatch.setenv("SRST_ST_FWD_CFG", str(request.param))
    monkeypatch.setenv("SRST_ST_BWD_CFG", str({0: 0, 1: 7, 2: 8, 3: 3, 4: 5, 10: 6}[request.param]))
    return request.param


@pytest.mark.parametrize("name", ["st_rand_2x24x36", "st_srlike_2x40x52", "st_rand_ragged_1x37x53"])
def test_emulated_kernels_match_oracle_on_golden_inputs(lib, tile_cfg, name):
    z = golden(name)
    taps = (z["g"], z["dg"], z["k"])
    out = emu_st(lib, z["sr"], z["hr"], taps, want_hr=True)
    ref = O.st_loss(z["sr"], z["hr"], taps=taps, want_hr_grad=True)
    assert rel_err(out["loss"], ref["loss"]) < 1e-5
    assert maxnorm_err(out["d_sr"], ref["d_sr"]) < 1e-4
    assert maxnorm_err(out["d_hr"], ref["d_hr"]) < 1e-4
    assert rel_err(out["loss"], z["loss"]) < 1e-5          # and the reference's own value
    assert np.all(out["ws"] == 0)                           # workspace handed back zeroed


@pytest.mark.parametrize("shape", [(1, 100, 152), (2, 96, 96), (1, 50, 203)])
def test_emulated_kernels_tile_seams(lib, tile_cfg, shape):
    rng = np.random.default_rng(sum(shape))
    sr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    hr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    taps = (*O.gaussian_taps(0.5, True), O.gaussian_taps(2.0))
    out = emu_st(lib, sr, hr, taps)
    ref = O.st_loss(sr, hr, taps=taps)
    assert rel_err(out["loss"], ref["loss"]) < 1e-5
    assert maxnorm_err(out["d_sr"], ref["d_sr"]) < 1e-4
    assert not np.isnan(out["d_sr"]).any()


def test_emulated_grad_out_scaling_and_nonorm(lib):
    z = golden("st_rand_2x24x36")
    taps = (z["g"], z["dg"], z["k"])
    a = emu_st(lib, z["sr"], z["hr"], taps, grad_out=1.0)
    b = emu_st(lib, z["sr"], z["hr"], taps, grad_out=-2.5)
    assert np.allclose(b["d_sr"], -2.5 * a["d_sr"], rtol=1e-6, atol=0)
    zn = golden("st_rand_nonorm_1x24x36")
    c = emu_st(lib, zn["sr"], zn["hr"], taps, normalize=False)
    assert rel_err(c["loss"], zn["loss"]) < 1e-5 and np.abs(c["d_sr"]).max() == 0.0


@pytest.mark.parametrize("sigma,rho", [(1.0, 2.5), (0.3, 1.0), (0.75, 1.5), (0.5, 3.0)])
def test_emulated_other_filter_radii(lib, sigma, rho):
    """Non-default sigma/rho run through the padded radius classes (odd radii get zero taps)."""
    rng = np.random.default_rng(int(10 * sigma + rho))
    sr = rng.random((1, 3, 40, 72), dtype=np.float32)
    hr = rng.random((1, 3, 40, 72), dtype=np.float32)
    taps = (*O.gaussian_taps(sigma, True), O.gaussian_taps(rho))
    out = emu_st(lib, sr, hr, taps)
    ref = O.st_loss(sr, hr, taps=taps)
    assert rel_err(out["loss"], ref["loss"]) < 1e-5
    assert maxnorm_err(out["d_sr"], ref["d_sr"]) < 1e-4


def test_emulated_golden_sigma1_rho25(lib):
    z = golden("st_rand_s1_r25_1x32x40")
    taps = (z["g"], z["dg"], z["k"])
    out = emu_st(lib, z["sr"], z["hr"], taps, want_hr=True)
    assert rel_err(out["loss"], z["loss"]) < 1e-5
    ref = O.st_loss(z["sr"], z["hr"], taps=taps, want_hr_grad=True)
    assert maxnorm_err(out["d_sr"], ref["d_sr"]) < 1e-4 and maxnorm_err(out["d_hr"], ref["d_hr"]) < 1e-4


def test_emulated_backward_without_saved_gray(lib):
    """gray = NULL: the backward re-reads RGB and converts (same result as the TMA gray-tile path)."""
    z = golden("st_srlike_2x40x52")
    taps = (z["g"], z["dg"], z["k"])
    a = emu_st(lib, z["sr"], z["hr"], taps, save_gray=True)
    b = emu_st(lib, z["sr"], z["hr"], taps, save_gray=False)
    assert np.allclose(a["d_sr"], b["d_sr"], rtol=1e-5, atol=1e-9)
    gray = 0.2989 * z["sr"][:, 0] + 0.587 * z["sr"][:, 1] + 0.114 * z["sr"][:, 2]
    assert np.allclose(a["gray_sr"], gray, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("shape", [(1, 3, 1, 1), (1, 3, 5, 7), (2, 3, 17, 4), (1, 3, 2, 130), (1, 3, 9, 9)])
def test_emulated_tiny_and_degenerate_shapes(lib, shape):
    """Images smaller than a tile, a single pixel, one-row strips: everything is halo."""
    rng = np.random.default_rng(sum(shape))
    sr = rng.random(shape, dtype=np.float32)
    hr = rng.random(shape, dtype=np.float32)
    z = golden("st_rand_2x24x36")
    taps = (z["g"], z["dg"], z["k"])
    out = emu_st(lib, sr, hr, taps, want_hr=True)
    ref = O.st_loss(sr, hr, taps=taps, want_hr_grad=True)
    assert rel_err(out["loss"], ref["loss"]) < 1e-5
    assert maxnorm_err(out["d_sr"], ref["d_sr"]) < 1e-4 and maxnorm_err(out["d_hr"], ref["d_hr"]) < 1e-4
