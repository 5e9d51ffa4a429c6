"""Filter taps for the structure tensor, computed once on the host.

The expressions are the reference's own (utils.py:194-208 ``get_gaussian_kernel``), evaluated with
torch on the CPU so the fp32 weights are bit-identical to what the reference builds (and moves to
the GPU) on every call; here they are cached per sigma and handed to the kernels as parameters.
"""
from __future__ import annotations

import ctypes
import functools

import numpy as np
import torch


def radius_of(sigma: float) -> int:
    return max(int(4 * sigma + 0.5), 1)  # utils.py:198


@functools.lru_cache(maxsize=64)
def gaussian_taps(sigma: float):
    """(g, dg) as contiguous float32 numpy arrays of length 2*radius+1."""
    radius = radius_of(sigma)
    x = torch.arange(-radius, radius + 1)
    sigma2 = (sigma * sigma) + 1e-12
    phi_x = torch.exp(-0.5 / sigma2 * x ** 2)
    phi_x = phi_x / phi_x.sum()
    dg = phi_x * -x / sigma2
    g_np = np.ascontiguousarray(phi_x.numpy(), dtype=np.float32)
    dg_np = np.ascontiguousarray(dg.numpy(), dtype=np.float32)
    g_np.setflags(write=False)
    dg_np.setflags(write=False)
    return g_np, dg_np


def as_c(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
