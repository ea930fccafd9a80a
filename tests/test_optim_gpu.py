"""hallucidet_b200.optim.FusedAdam (clip + Adam + bf16 re-pack in one pass) against the reference's optimizer tail:
``clip_grad_value_(0.5)`` + ``torch.optim.Adam(lr=1e-4)`` (train_hallucidet.py:429-435,498-499) on identical gradients."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def _unet(seed=3):
    from hallucidet_b200.unet import Unet
    torch.manual_seed(seed)
    m = Unet("resnet34", classes=3)
    m.segmentation_head[-1] = torch.nn.Sigmoid()
    return m.cuda().train()


@pytest.mark.parametrize("grad_scale", [1.0, 0.25])
def test_fused_adam_matches_torch_adam(grad_scale):
    from hallucidet_b200.optim import FusedAdam
    a = _unet()
    b = copy.deepcopy(a)
    opt_a = torch.optim.Adam(a.parameters(), lr=1e-4)
    opt_b = FusedAdam(b, lr=1e-4, clip_value=0.5, grad_scale=grad_scale)
    gen = torch.Generator().manual_seed(1)
    for step in range(3):
        x = torch.rand(2, 1, 64, 96, generator=gen).cuda().expand(-1, 3, -1, -1)
        for m in (a, b):
            m.zero_grad(set_to_none=True)
            (m(x) * 37.0).square().sum().backward()           # large loss: many gradient entries exceed the clip value
        torch.cuda.synchronize()
        # identical gradients on both sides (a backward is only reproducible to ~1e-2 run to run, see test_modules_gpu)
        eng_b = next(e for e in b._engines.values() if e.training)
        ga = torch.cat([p.grad.flatten() for p in a.parameters()])
        eng_b.flat_grad.copy_(ga)
        assert float((ga.abs() > 0.5 / grad_scale).float().mean()) > 0.001           # the clip is exercised
        for p in a.parameters():
            p.grad.mul_(grad_scale)
        torch.nn.utils.clip_grad_value_(a.parameters(), 0.5)
        opt_a.step()
        opt_b.step()
        torch.cuda.synchronize()
        for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            assert torch.allclose(pa, pb, rtol=1e-5, atol=2e-7), (step, n, float((pa - pb).abs().max()))
        sa, sb = opt_a.state_dict()["state"], opt_b.state_dict()["state"]
        assert len(sa) == len(sb)
        for k in (0, len(sa) // 2, len(sa) - 1):
            assert torch.allclose(sa[k]["exp_avg"], sb[k]["exp_avg"], rtol=1e-5, atol=1e-8)
            assert torch.allclose(sa[k]["exp_avg_sq"], sb[k]["exp_avg_sq"], rtol=1e-5, atol=1e-10)
            assert float(sb[k]["step"]) == step + 1
    # the bf16 operands written by the fused pass == a fresh pack of the updated masters (the next forward skips its pack)
    xe = torch.rand(2, 1, 64, 96, generator=gen).cuda().expand(-1, 3, -1, -1)
    with torch.no_grad():
        h1 = b(xe).clone()
        assert eng_b._packed_key == eng_b._version_key()
        eng_b._packed_key = None                                  # force the stand-alone pack launch
        h2 = b(xe).clone()
    assert torch.equal(h1, h2)
    # eval-mode engine notices that the parameters moved (raw-pointer updates do not bump tensor versions)
    b.eval()
    with torch.no_grad():
        e1 = b(xe).clone()
    b.train()
    b.zero_grad(set_to_none=True)
    (b(xe) * 37.0).square().sum().backward()
    opt_b.step()
    b.eval()
    with torch.no_grad():
        e2 = b(xe).clone()
    assert not torch.equal(e1, e2)


def test_fused_adam_state_dict_roundtrip_and_lr_schedule():
    from hallucidet_b200.optim import FusedAdam
    m = _unet()
    opt = FusedAdam(m, lr=1e-4, clip_value=0.5)
    x = torch.rand(2, 3, 64, 64).cuda()
    m(x).square().sum().backward()
    opt.step()
    sd = copy.deepcopy(opt.state_dict())
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min")       # train_hallucidet.py:437-445
    for _ in range(12):
        sched.step(1.0)
    assert opt.param_groups[0]["lr"] < 1e-4
    opt2 = FusedAdam(m, lr=1e-4, clip_value=0.5)
    opt2.load_state_dict(sd)
    assert opt2._steps == 1
    assert torch.equal(opt2.state_dict()["state"][5]["exp_avg"], sd["state"][5]["exp_avg"])
    m.zero_grad(set_to_none=True)
    m(x).square().sum().backward()
    opt2.step()
    torch.cuda.synchronize()
    assert float(opt2.state_dict()["state"][0]["step"]) == 2
    with pytest.raises(NotImplementedError):
        FusedAdam(m, lr=1e-4, weight_decay=0.1)
