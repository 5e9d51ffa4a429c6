"""CPU: execute the real Best-Buddy kernel sources under the test-only host emulation and compare
indices bit-exactly with the C oracle (same fp32 operation order) and with the reference's golden
indices."""
import numpy as np
import pytest

from oracle import bb_oracle as O
from tests.helpers import emu_bb, emu_lib, golden, golden_names, maxnorm_err, rel_err


@pytest.fixture(scope="module")
def lib():
    return emu_lib()


@pytest.mark.parametrize("name", golden_names("bb_"))
@pytest.mark.parametrize("own_pyramid", [False, True])
def test_emulated_bb_matches_oracle_and_reference(lib, name, own_pyramid):
    z = golden(name)
    a, b = float(z["alpha"]), float(z["beta"])
    crit = 0 if str(z["criterion"]) == "l1" else 1
    gt2, gt4 = (None, None) if own_pyramid else (z["hr2"], z["hr4"])
    out = emu_bb(lib, z["sr"], z["hr"], gt2, gt4, a, b, crit)
    orc = O.bb_forward_c(z["sr"], z["hr"], gt2, gt4, a, b, str(z["criterion"]))
    assert np.array_equal(out["idx"], orc["idx"]), "indices must be bit-exact vs the C oracle"
    assert np.array_equal(out["idx"], z["ind"]), "and equal the reference's torch.min indices"
    assert rel_err(out["loss"], z["loss"]) < 1e-5 or abs(out["loss"] - float(z["loss"])) < 1e-9
    if np.abs(z["d_sr"]).max() == 0:
        assert np.abs(out["d_sr"]).max() == 0
    else:
        assert maxnorm_err(out["d_sr"], z["d_sr"]) < 1e-5


def test_emulated_pyramid_kernel_matches_oracle(lib):
    import ctypes
    z = golden("bb_rand_1x48x36")
    hr = np.ascontiguousarray(z["hr"], np.float32)
    o2 = np.empty_like(z["hr2"]); o4 = np.empty_like(z["hr4"])
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert lib.srst_bb_pyramid(p(hr), 1, 48, 36, p(o2), p(o4), None) == 0
    r2, r4 = O.pyramid_c(hr)
    assert np.array_equal(o2, r2) and np.array_equal(o4, r4)
    assert np.abs(o2 - z["hr2"]).max() < 5e-7


def test_emulated_bb_ragged_shape_and_ties(lib):
    """H, W not multiples of 3/12 (floor semantics of unfold/interpolate) and exact ties: a constant
    image makes every candidate equal, so index 0 must win everywhere (torch.min rule)."""
    rng = np.random.default_rng(3)
    sr = rng.random((1, 3, 26, 31), dtype=np.float32)
    gt = rng.random((1, 3, 26, 31), dtype=np.float32)
    out = emu_bb(lib, sr, gt)
    orc = O.bb_forward_c(sr, gt)
    assert np.array_equal(out["idx"], orc["idx"]) and rel_err(out["loss"], orc["loss"]) < 1e-6
    assert np.all(out["d_sr"][:, :, 24:, :] == 0) and np.all(out["d_sr"][:, :, :, 30:] == 0)
    flat = np.full((1, 3, 24, 24), 0.25, np.float32)
    out = emu_bb(lib, flat, flat)
    assert np.all(out["idx"] == 0) and out["loss"] == 0.0


@pytest.mark.parametrize("alpha,beta", [(0.0, 1.0), (1.0, 0.0), (0.7, 1.3), (-0.25, 1.0)])
def test_emulated_search_filter_is_exact_for_any_weights(lib, alpha, beta):
    """The search kernel discards candidates with a single-dot lower bound and re-scores survivors
    exactly: for zero, unequal and even negative weights the indices must still equal exhaustive
    exact scoring (the C oracle), on smooth SR-like data (co-located patch usually wins) and on noise."""
    rng = np.random.default_rng(17)
    gt = rng.random((1, 3, 36, 48), dtype=np.float32)
    for sr in (np.clip(gt + 0.05 * rng.standard_normal(gt.shape).astype(np.float32), 0, 1),
               rng.random(gt.shape, dtype=np.float32)):
        out = emu_bb(lib, sr, gt, alpha=alpha, beta=beta)
        orc = O.bb_forward_c(sr, gt, alpha=alpha, beta=beta)
        assert np.array_equal(out["idx"], orc["idx"])
        assert rel_err(out["loss"], orc["loss"]) < 1e-6


def test_emulated_search_many_near_ties(lib):
    """Quantised 2-level images make thousands of exactly tied and nearly tied scores: the filter must
    keep every co-minimal candidate so that the lowest index wins (torch.min rule)."""
    rng = np.random.default_rng(23)
    gt = (rng.random((1, 3, 48, 48)) > 0.5).astype(np.float32)
    sr = (rng.random((1, 3, 48, 48)) > 0.5).astype(np.float32)
    out = emu_bb(lib, sr, gt)
    orc = O.bb_forward_c(sr, gt)
    assert np.array_equal(out["idx"], orc["idx"])


@pytest.mark.parametrize("mode", ["patch", "gram", "pst"])
@pytest.mark.parametrize("where", ["sr", "gt"])
def test_emulated_nan_input_gives_nan_loss_and_in_range_indices(lib, mode, where):
    """A NaN pixel must propagate like in the reference (torch.clamp keeps NaN, torch.min returns the
    first NaN: utils.py:187, loss.py:135) -- a NaN loss, every index inside [0, M), no out-of-bounds read
    in the loss / backward kernels that consume the indices (the search used to emit 0x7fffffff)."""
    rng = np.random.default_rng(5)
    gt = rng.random((1, 3, 24, 24), dtype=np.float32)
    sr = np.clip(gt + 0.05 * rng.standard_normal(gt.shape).astype(np.float32), 0, 1)
    (sr if where == "sr" else gt)[0, 1, 7, 8] = np.nan
    taps = None
    if mode == "pst":
        from oracle import st_oracle as S
        taps = (*S.gaussian_taps(0.5, True), S.gaussian_taps(2.0))
    out = emu_bb(lib, sr, gt, mode=mode, taps=taps)
    N = 8 * 8
    M = N + 4 * 4 + 2 * 2
    assert out["idx"].min() >= 0 and out["idx"].max() < M
    assert np.isnan(out["loss"])
    if mode == "patch":
        orc = O.bb_forward_c(sr, gt)
        assert np.array_equal(out["idx"], orc["idx"]), "same torch.min NaN order as the oracle"
        if where == "sr":
            assert out["idx"][0, 2 * 8 + 2] == 0          # the query holding the NaN: every score NaN -> index 0
        else:
            want = np.full(N, 2 * 8 + 2)                  # the level-0 candidate holding the NaN wins every row ...
            want[2 * 8 + 2] = 0                           # ... except its own query (g_i is NaN: every score NaN)
            assert np.array_equal(out["idx"][0], want)
