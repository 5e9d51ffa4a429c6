"""GPU: the modules drop into the reference's training-loop contract (train.py:130-144,
warmup.py:86-96, config.py:71-93,122-125): a name->module registry with weights, `loss*weight`
accumulation, `.item()` after every criterion, one backward over the sum, an optimizer step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Cfg:
    """Just the plugin surface of reference config.py (G_LOSS.CRITERIONS / CRITERION_WEIGHTS)."""

    def __init__(self):
        self.CRITERIONS = {"Pixel": torch.nn.MSELoss()}
        self.CRITERION_WEIGHTS = {"Pixel": 1.0}

    def add_g_criterion(self, name, module, weight):  # config.py:122-125
        self.CRITERIONS[name] = module
        self.CRITERION_WEIGHTS[name] = weight


def test_losses_register_and_train_like_the_reference_loop():
    from srgan_st_b200 import BestBuddyLoss, StructureTensorLoss
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    cfg = _Cfg()
    cfg.add_g_criterion("ST", StructureTensorLoss().to(dev), 1 / 3)        # config.py:80 weight
    cfg.add_g_criterion("BestBuddy", BestBuddyLoss().to(dev), 50.0)        # config.py:82 weight
    gen = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.PReLU(),
                              torch.nn.Conv2d(16, 3, 3, padding=1)).to(dev)
    opt = torch.optim.Adam(gen.parameters(), lr=1e-3)
    gt = (torch.randint(0, 256, (8, 3, 96, 96), device=dev).float() / 255)
    lr = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)
    history = []
    for _ in range(6):
        sr = gen(lr).clamp(0, 1)
        g_loss = torch.tensor(0.0, device=dev)
        loss_values = {}
        for name, criterion in cfg.CRITERIONS.items():
            loss = criterion(sr, gt)
            weight = cfg.CRITERION_WEIGHTS[name]
            g_loss = g_loss + (loss * weight)
            loss_values[name] = (loss * weight).item()      # the per-criterion sync of train.py:141
        opt.zero_grad()
        g_loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in gen.parameters())
        opt.step()
        history.append(g_loss.item())
        assert set(loss_values) == {"Pixel", "ST", "BestBuddy"}
        assert all(v == v and v >= 0 for v in loss_values.values())
    assert history[-1] < history[0]


def test_every_loss_module_trains_in_the_loop_and_under_a_cuda_graph():
    """All five modules (BestBuddy twice: default and a non-default patch geometry) registered at once (incl. the fused ST+Pixel criterion replacing "Pixel" + "ST"), then the
    whole generator step replayed from ONE CUDA graph without any per-criterion .item(): the
    loss objects neither allocate outside torch's caching allocator nor synchronise."""
    from srgan_st_b200 import (BestBuddyLoss, GramLoss, PatchwiseStructureTensorLoss, StructureTensorLoss,
                               StructureTensorPixelLoss)
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    crits = {"ST+Pixel": (StructureTensorPixelLoss(st_weight=1 / 3, pixel_weight=1.0), 1.0),
             "BestBuddy": (BestBuddyLoss(), 50.0), "Gram": (GramLoss(), 10.0),
             "PatchST": (PatchwiseStructureTensorLoss(), 1.0), "ST": (StructureTensorLoss(rho=1.0), 0.1),
             "BestBuddy4": (BestBuddyLoss(ksize=4, pad=1, stride=2), 20.0)}   # overlapping patches: the all-pairs path
    gen = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.PReLU(),
                              torch.nn.Conv2d(8, 3, 3, padding=1)).to(dev)
    opt = torch.optim.SGD(gen.parameters(), lr=1e-3)
    gt = (torch.randint(0, 256, (4, 3, 48, 48), device=dev).float() / 255)
    lr = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)

    def g_step():
        sr = gen(lr).clamp(0, 1)
        total = sum(w * m(sr, gt) for m, w in crits.values())
        opt.zero_grad(set_to_none=False)
        total.backward()
        return total

    eager = [g_step().item() for _ in range(3)]          # warm-up on the default stream, also fills .grad
    assert all(torch.isfinite(p.grad).all() for p in gen.parameters())
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            g_step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        static_total = g_step()
    graph.replay()
    torch.cuda.synchronize()
    replayed = static_total.item()
    grads = [p.grad.clone() for p in gen.parameters()]
    ref = g_step()                                        # same weights, eager
    torch.cuda.synchronize()
    assert abs(replayed - ref.item()) <= 1e-5 * abs(ref.item())
    for a, p in zip(grads, gen.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-4, atol=1e-7)
    assert eager[0] == eager[0]


def test_non_contiguous_and_sliced_inputs_are_accepted():
    from srgan_st_b200 import StructureTensorLoss
    x = torch.rand(4, 3, 40, 44, device="cuda").requires_grad_(True)
    gt = torch.rand(4, 3, 40, 44, device="cuda")
    m = StructureTensorLoss()
    a = m(x[1:3], gt[1:3])                       # offset base pointer (still 16-byte aligned or not)
    b = m(x[1:3].clone(), gt[1:3].clone())
    assert torch.allclose(a, b, rtol=1e-6)
    xt = torch.rand(2, 40, 44, 3, device="cuda").permute(0, 3, 1, 2)   # channels-last view
    gtt = torch.rand(2, 40, 44, 3, device="cuda").permute(0, 3, 1, 2)
    c = m(xt, gtt)
    d = m(xt.contiguous(), gtt.contiguous())
    assert torch.allclose(c, d, rtol=1e-6)
