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


@pytest.fixture(params=[(0, 0), (1, 1), (2, 0), (3, 1), (4, 0)], ids=lambda p: f"fwd{p[0]}-bwd{p[1]}")
def tile_cfg(request, lib):
    """Every compiled forward shape (0-2 tiled, 3-4 row-marching), paired with a backward shape (srst_st_force_cfg)."""
    assert lib.srst_st_num_cfgs(0) == 5 and lib.srst_st_num_cfgs(1) == 2
    assert lib.srst_st_force_cfg(*request.param) == 0
    yield request.param
    lib.srst_st_force_cfg(-1, -1)


@pytest.mark.parametrize("cfg", [3, 4])
@pytest.mark.parametrize("chunk_blocks", [1, 2, 3])
@pytest.mark.parametrize("shape", [(1, 100, 152), (2, 96, 96), (1, 37, 52), (1, 52, 204)])
def test_emulated_marching_forward_row_chunks(lib, cfg, chunk_blocks, shape):
    """The row-marching forward cut into chunks of 16, 32, 48 rows: chunk seams, the warm-up block above a chunk, the
    ragged last block and strips wider than the image must all reproduce the oracle (and the saved Ix, Iy planes)."""
    assert lib.srst_st_force_cfg(cfg, -1) == 0 and lib.srst_st_force_chunk_blocks(chunk_blocks) == 0
    try:
        rng = np.random.default_rng(sum(shape) + chunk_blocks)
        sr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
        hr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
        taps = (*O.gaussian_taps(0.5, True), O.gaussian_taps(2.0))
        out = emu_st(lib, sr, hr, taps, want_hr=True)
        ref = O.st_loss(sr, hr, taps=taps, want_hr_grad=True)
        assert rel_err(out["loss"], ref["loss"]) < 1e-5
        assert maxnorm_err(out["d_sr"], ref["d_sr"]) < 1e-4 and maxnorm_err(out["d_hr"], ref["d_hr"]) < 1e-4
        assert not np.isnan(out["ixy_sr"]).any() and np.all(out["ws"] == 0)
    finally:
        lib.srst_st_force_cfg(-1, -1)
        lib.srst_st_force_chunk_blocks(0)


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


def test_emulated_saved_gradient_planes(lib):
    """The ixy planes the forward saves are the oracle's Ix, Iy, row-pair interleaved, zero in the padding row of an
    odd-height image -- and fully written (no NaN left from the poison fill)."""
    z = golden("st_rand_ragged_1x37x53")
    taps = (z["g"], z["dg"], z["k"])
    out = emu_st(lib, z["sr"], z["hr"], taps)
    gray = (0.2989 * z["sr"][:, 0] + 0.587 * z["sr"][:, 1]) + 0.114 * z["sr"][:, 2]
    Ix = O._corr_axis(O._corr_axis(gray.astype(np.float64), z["dg"].astype(np.float64), 1), z["g"].astype(np.float64), 2)
    Iy = O._corr_axis(O._corr_axis(gray.astype(np.float64), z["g"].astype(np.float64), 1), z["dg"].astype(np.float64), 2)
    ixy = out["ixy_sr"]                                  # [B, 2, Hp, W, 2]
    H = gray.shape[1]
    got_x = ixy[:, 0].transpose(0, 1, 3, 2).reshape(1, -1, gray.shape[2])[:, :H]   # rows 2p, 2p+1 de-interleaved
    got_y = ixy[:, 1].transpose(0, 1, 3, 2).reshape(1, -1, gray.shape[2])[:, :H]
    assert not np.isnan(ixy).any()
    assert np.allclose(got_x, Ix, rtol=1e-5, atol=1e-6) and np.allclose(got_y, Iy, rtol=1e-5, atol=1e-6)
    assert np.all(ixy[:, :, -1, :, 1] == 0)             # H = 37 is odd: the pair partner of the last row


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


@pytest.mark.parametrize("name", ["st_rand_s15_r4_1x40x52", "st_srlike_s2_r5_2x33x45"])
def test_emulated_generic_radius_matches_reference(lib, name):
    """Radii beyond the compiled classes: the generic-radius kernels (st_generic.cuh) against the reference's outputs
    and the fp64 oracle, both gradients, through srst_st_workspace_bytes_r / srst_st_backward_ws."""
    z = golden(name)
    taps = (z["g"], z["dg"], z["k"])
    assert lib.srst_st_supported(len(z["g"]) // 2, len(z["k"]) // 2) == 2
    out = emu_st(lib, z["sr"], z["hr"], taps, want_hr=True)
    ref = O.st_loss(z["sr"], z["hr"], taps=taps, want_hr_grad=True)
    assert rel_err(out["loss"], z["loss"]) < 1e-5 and rel_err(out["loss"], ref["loss"]) < 1e-5
    for ours, orc, refg in ((out["d_sr"], ref["d_sr"], z["d_sr"]), (out["d_hr"], ref["d_hr"], z["d_hr"])):
        assert maxnorm_err(ours, orc) < 1e-4
        assert maxnorm_err(ours, refg) < 1e-4 + maxnorm_err(refg, orc)
    assert np.all(out["ws"] == 0)   # ticket header handed back zeroed


@pytest.mark.parametrize("sigma,rho,shape", [(2.0, 2.0, (1, 20, 28)), (0.5, 4.0, (2, 17, 19)), (3.0, 8.0, (1, 9, 13))])
def test_emulated_generic_radius_one_sided_and_larger_than_image(lib, sigma, rho, shape):
    rng = np.random.default_rng(int(10 * sigma + rho))
    sr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    hr = rng.random((shape[0], 3, shape[1], shape[2]), dtype=np.float32)
    taps = (*O.gaussian_taps(sigma, True), O.gaussian_taps(rho))
    out = emu_st(lib, sr, hr, taps, want_hr=True, grad_out=-0.5)
    ref = O.st_loss(sr, hr, taps=taps, want_hr_grad=True)
    assert rel_err(out["loss"], ref["loss"]) < 1e-5
    assert maxnorm_err(out["d_sr"], -0.5 * ref["d_sr"]) < 1e-4 and maxnorm_err(out["d_hr"], -0.5 * ref["d_hr"]) < 1e-4


def test_generic_radius_limits_and_workspace_sizes(lib):
    assert lib.srst_st_supported(2, 8) == 1 and lib.srst_st_supported(4, 12) == 1
    assert lib.srst_st_supported(5, 8) == 2 and lib.srst_st_supported(2, 13) == 2 and lib.srst_st_supported(64, 64) == 2
    assert lib.srst_st_supported(65, 8) == 0 and lib.srst_st_supported(2, 65) == 0 and lib.srst_st_supported(0, 8) == 0
    base = lib.srst_st_workspace_bytes(2, 30, 40)
    assert lib.srst_st_workspace_bytes_r(2, 30, 40, 2, 8) == base and lib.srst_st_backward_workspace_bytes(2, 30, 40, 2, 8) == 0
    assert lib.srst_st_workspace_bytes_r(2, 30, 40, 6, 16) >= base + 11 * 2 * 30 * 40 * 4
    assert lib.srst_st_backward_workspace_bytes(2, 30, 40, 6, 16) >= 5 * 2 * 30 * 40 * 4
    assert lib.srst_st_workspace_bytes_r(2, 30, 40, 65, 16) == 0
