"""CPU: host-side logic of the package (taps, module surface, argument checks)."""
import numpy as np
import pytest
import torch

from tests.helpers import golden


def test_host_taps_are_bit_identical_to_the_reference():
    from srgan_st_b200 import taps
    for name, sigma, rho in (("st_rand_2x24x36", 0.5, 2.0), ("st_rand_s1_r25_1x32x40", 1.0, 2.5)):
        z = golden(name)
        g, dg = taps.gaussian_taps(sigma)
        k, _ = taps.gaussian_taps(rho)
        assert np.array_equal(g, z["g"]) and np.array_equal(dg, z["dg"]) and np.array_equal(k, z["k"])
    assert taps.radius_of(0.5) == 2 and taps.radius_of(2.0) == 8 and taps.radius_of(0.1) == 1


def test_module_surface_matches_reference():
    import srgan_st_b200 as pkg
    m = pkg.StructureTensorLoss()
    assert (m.sigma, m.rho, m.normalize) == (0.5, 2.0, True)
    m2 = pkg.StructureTensorLoss(sigma=1.0, rho=2.5, normalize=False)
    assert (m2.sigma, m2.rho, m2.normalize) == (1.0, 2.5, False)
    assert isinstance(m, torch.nn.Module) and len(list(m.parameters())) == 0


def test_cpu_tensors_raise_instead_of_falling_back():
    import srgan_st_b200 as pkg
    a = torch.rand(1, 3, 16, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.StructureTensorLoss()(a, a)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from srgan_st_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.SrstError, match="no CPU or PyTorch fallback"):
        _cabi.lib()


def test_best_buddy_module_surface_matches_reference():
    import srgan_st_b200 as pkg
    m = pkg.BestBuddyLoss()
    assert (m.alpha, m.beta, m.ksize, m.pad, m.stride, m.dist_norm) == (1.0, 1.0, 3, 0, 3, "l2")
    assert isinstance(m.criterion, torch.nn.L1Loss)
    assert isinstance(pkg.BestBuddyLoss(criterion="mse").criterion, torch.nn.MSELoss)
    with pytest.raises(NotImplementedError):
        pkg.BestBuddyLoss(criterion="huber")
