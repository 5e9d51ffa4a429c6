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
