"""CPU, world_size 2 (gloo): the data-parallel plumbing around the loss path -- batch sharding and
the flat gradient bucket with the loss in its tail -- reproduces the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 3, 3, padding=1))


def _fake_loss(sr, gt):
    # stands in for criterion(sr, gt): a per-image mean, like the ST loss (the kernels need a GPU)
    return ((sr - gt) ** 2).mean()


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from srgan_st_b200.dist import FlatGradBucket, shard_batch
    g = _model()
    bucket = FlatGradBucket(g.parameters())
    torch.manual_seed(1)
    lr, gt = torch.rand(8, 3, 12, 12), torch.rand(8, 3, 12, 12)
    bucket.zero()
    loss = _fake_loss(g(shard_batch(lr, rank, world)), shard_batch(gt, rank, world))
    loss.backward()                      # writes into the bucket views
    bucket.set_loss(loss)
    mean_loss = bucket.all_reduce_mean()
    if rank == 0:
        torch.save({"loss": mean_loss.clone(), "flat": bucket.flat.clone()}, out)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    g = _model()
    torch.manual_seed(1)
    lr, gt = torch.rand(8, 3, 12, 12), torch.rand(8, 3, 12, 12)
    loss = _fake_loss(g(lr), gt)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in g.parameters()])
    assert torch.allclose(got["loss"], loss.detach(), rtol=1e-6, atol=1e-8)
    assert torch.allclose(got["flat"][:-1], ref, rtol=1e-5, atol=1e-8)
    assert torch.allclose(got["flat"][-1], loss.detach(), rtol=1e-6, atol=1e-8)


def test_shard_batch_requires_even_split():
    from srgan_st_b200.dist import shard_batch
    t = torch.arange(12).reshape(6, 2)
    assert torch.equal(shard_batch(t, 1, 3), t[2:4])
    with pytest.raises(ValueError):
        shard_batch(t, 0, 4)


def test_bucket_grads_are_views():
    from srgan_st_b200.dist import FlatGradBucket
    g = _model()
    b = FlatGradBucket(g.parameters())
    n = sum(p.numel() for p in g.parameters())
    assert b.flat.numel() == n + 1 and b.nbytes == 4 * (n + 1)
    g(torch.rand(1, 3, 8, 8)).sum().backward()
    assert b.flat[:-1].abs().sum() > 0          # backward wrote through the views
    assert all(p.grad.data_ptr() >= b.flat.data_ptr() for p in g.parameters())


def test_bucket_survives_zero_grad_set_to_none():
    """optimizer.zero_grad() / module.zero_grad() (set_to_none=True by default) drop the views; the bucket must
    notice and re-attach, otherwise the collective reduces stale zeros and ranks silently diverge."""
    from srgan_st_b200.dist import FlatGradBucket
    g = _model()
    b = FlatGradBucket(g.parameters())
    opt = torch.optim.SGD(g.parameters(), lr=0.1)
    x = torch.rand(2, 3, 8, 8)
    g(x).sum().backward()
    want = torch.cat([p.grad.reshape(-1) for p in g.parameters()]).clone()
    opt.zero_grad()                                   # set_to_none=True: every p.grad is None now
    assert all(p.grad is None for p in g.parameters())
    g(x).sum().backward()                             # fresh grad tensors OUTSIDE the bucket
    assert any(p.grad.data_ptr() < b.flat.data_ptr() or p.grad.data_ptr() >= b.flat.data_ptr() + b.nbytes
               for p in g.parameters())
    b.all_reduce_mean()                               # world size 1: no collective, but the repair runs
    assert torch.allclose(b.flat[:-1], want)          # the stray gradients were copied in
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(b.params, b._views))
    g.zero_grad()                                     # again dropped ...
    b.zero()                                          # ... and re-attached, zeroed
    assert all(p.grad is not None and p.grad.abs().sum() == 0 for p in g.parameters())
    g(x).sum().backward()
    assert torch.allclose(b.flat[:-1], want)
