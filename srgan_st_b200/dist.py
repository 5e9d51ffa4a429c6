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
    collective with no packing copy.

    ``optimizer.zero_grad()`` / ``module.zero_grad()`` default to ``set_to_none=True``, which DROPS the
    views: the next backward would then allocate fresh grad tensors outside the bucket and the collective
    would reduce stale zeros.  Use ``bucket.zero()`` (or ``zero_grad(set_to_none=False)``); as a safety
    net ``zero()`` and ``all_reduce_mean()`` re-attach a view whenever a parameter's ``.grad`` no longer
    aliases the bucket (copying a stray gradient in first, so nothing computed is lost)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n + 1, dtype=torch.float32, device=dev)
        self._views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatGradBucket expects fp32 parameters on one device")
            v = self.flat[off:off + p.numel()].view_as(p)
            self._views.append(v)
            p.grad = v
            off += p.numel()
        self.loss_slot = self.flat[n:n + 1]
        # side stream + event for the overlapped collective (all_reduce_mean_async)
        self._comm_stream = None
        self._done = None

    def reattach(self) -> int:
        """Make every ``param.grad`` a view of the bucket again; returns how many had to be repaired."""
        fixed = 0
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is None:
                v.zero_()
                p.grad = v
                fixed += 1
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)          # a gradient computed outside the bucket: bring it in
                p.grad = v
                fixed += 1
        return fixed

    def zero(self) -> None:
        self.flat.zero_()
        self.reattach()

    def set_loss(self, loss: torch.Tensor) -> None:
        self.loss_slot.copy_(loss.detach().reshape(1))

    def all_reduce_mean(self, group=None) -> torch.Tensor:
        """Sum over ranks, divide by the world size; returns the mean loss (a view, no sync)."""
        self.reattach()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)   # mean inside NCCL: no extra kernel
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        return self.loss_slot[0]

    def all_reduce_mean_async(self, group=None) -> None:
        """The same collective on a side stream: it starts once everything enqueued so far on the current
        stream (the backward pass that filled the bucket) has finished and overlaps whatever the caller
        enqueues next (the next step's H2D copies and loss kernels).  ``wait()`` makes the current stream
        wait for it -- call it before the optimizer step / before reading ``loss_slot``."""
        self.reattach()
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.flat.device)
            self._done = torch.cuda.Event()
        cur = torch.cuda.current_stream(self.flat.device)
        self._comm_stream.wait_stream(cur)
        with torch.cuda.stream(self._comm_stream):
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            self._done.record(self._comm_stream)
        self.flat.record_stream(self._comm_stream)

    def wait(self) -> torch.Tensor:
        """Order the current stream after the last asynchronous collective; returns the mean loss view."""
        if self._done is not None:
            torch.cuda.current_stream(self.flat.device).wait_event(self._done)
        return self.loss_slot[0]

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4
