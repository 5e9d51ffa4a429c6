"""Data-parallel plumbing for the loss hot path.

The two losses shard over images: every rank evaluates its slice of the batch with the fused
kernels and no collective touches the loss data path.  The only exchange of a training step is one
NCCL all-reduce over NVLink of a flat bucket holding the generator gradients with the scalar loss
appended as the last element (so the loss costs no collective of its own).  The reference is
single-GPU (config.py:17 ``cuda:0``); mean-of-means equals its global ``.mean()`` because every rank
gets the same number of equally sized images.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank's contiguous slice of the batch dimension; the batch must divide evenly so that the
    mean of per-rank means equals the reference's global mean (loss.py:413)."""
    if t.shape[0] % world != 0:
        raise ValueError(f"batch {t.shape[0]} does not divide evenly over {world} ranks")
    per = t.shape[0] // world
    return t[rank * per:(rank + 1) * per]


class FlatGradBucket:
    """One flat fp32 buffer: [grad of every parameter ... | loss].  ``param.grad`` tensors are views
    into the buffer, so backward writes straight into it and ``all_reduce_mean`` is a single
    collective with no packing copy."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n + 1, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatGradBucket expects fp32 parameters on one device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.loss_slot = self.flat[n:n + 1]

    def zero(self) -> None:
        self.flat.zero_()

    def set_loss(self, loss: torch.Tensor) -> None:
        self.loss_slot.copy_(loss.detach().reshape(1))

    def all_reduce_mean(self, group=None) -> torch.Tensor:
        """Sum over ranks, divide by the world size; returns the mean loss (a view, no sync)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.loss_slot[0]

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4
