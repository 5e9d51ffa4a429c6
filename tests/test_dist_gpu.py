"""GPU, world_size 2 (NCCL): the real StructureTensorLoss on batch shards, gradients through FlatGradBucket --
the mean of the per-rank losses equals the single-GPU loss on the concatenated batch, and the all-reduced
generator gradient equals the single-GPU gradient (images are independent, the only exchange is the
gradient bucket with the loss in its tail).  Needs two visible GPUs: skipped on a 1-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gen():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.PReLU(), torch.nn.Conv2d(8, 3, 3, padding=1))


def _data():
    g = torch.Generator().manual_seed(1)
    gt = torch.randint(0, 256, (8, 3, 96, 96), generator=g).float() / 255
    lr = (gt + 0.1 * torch.randn(8, 3, 96, 96, generator=g)).clamp(0, 1)
    return lr, gt


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from srgan_st_b200 import StructureTensorLoss
    from srgan_st_b200.dist import FlatGradBucket, shard_batch
    gen = _gen().to(dev)
    bucket = FlatGradBucket(gen.parameters())
    lr, gt = _data()
    crit = StructureTensorLoss()
    for mode in ("sync", "async"):
        bucket.zero()
        sr = gen(shard_batch(lr, rank, world).to(dev)).clamp(0, 1)
        loss = crit(sr, shard_batch(gt, rank, world).to(dev))
        loss.backward()
        bucket.set_loss(loss)
        if mode == "sync":
            mean_loss = bucket.all_reduce_mean()
        else:
            bucket.all_reduce_mean_async()
            mean_loss = bucket.wait()
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({"loss": mean_loss.cpu().clone(), "flat": bucket.flat.cpu().clone()}, out + "." + mode)
    dist.destroy_process_group()


def test_two_rank_st_loss_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    from srgan_st_b200 import StructureTensorLoss
    dev = torch.device("cuda:0")
    gen = _gen().to(dev)
    lr, gt = _data()
    loss = StructureTensorLoss()(gen(lr.to(dev)).clamp(0, 1), gt.to(dev))
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in gen.parameters()]).cpu()
    for mode in ("sync", "async"):
        got = torch.load(out + "." + mode)
        assert abs(got["loss"].item() - loss.item()) <= 1e-6 * abs(loss.item()), mode
        assert torch.allclose(got["flat"][-1], loss.detach().cpu(), rtol=1e-6, atol=0)
        err = (got["flat"][:-1] - ref).abs().max().item() / ref.abs().max().item()
        assert err < 1e-5, (mode, err)
