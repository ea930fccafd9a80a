"""Module-level parity on the GPU: the B200 U-Net / frozen backbone / transform / assembled train step against
the oracle (oracle/*.py) on the same seeded inputs and weights.

Two oracles are used (SURVEY.md section 7 "Hard parts"):
  * fp32 oracle            -- the reference restatement itself: gates the LOSS (<= 1 % relative) and reports hal error;
  * bf16-storage oracle    -- the same fp32 restatement with values rounded to bf16 at the points where the
    kernels store bf16 (conv outputs, activations, gradients; weights bf16) -- identical rounding on both
    sides, so only accumulation order differs: gates hal (max-abs <= 2e-2) and gradients (cosine >= 0.999).
"""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def flat(d, keys):
    return torch.cat([d[k].double().flatten() for k in keys])


def make_unet(seed=123):
    from hallucidet_b200.unet import Unet
    torch.manual_seed(seed)
    m = Unet("resnet34", encoder_depth=5, encoder_weights=None, decoder_attention_type=None, in_channels=3, classes=3)
    m.segmentation_head[-1] = torch.nn.Sigmoid()
    return m.cuda()


def oracle_state(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def test_unet_train_forward_backward(golden_dir):
    from oracle import unet as ou
    g = torch.load(os.path.join(golden_dir, "unet_small.pt"), weights_only=False)
    m = make_unet().train()
    state = oracle_state(m)
    x = g["ir"].repeat(1, 3, 1, 1).cuda()
    gw = torch.linspace(-1, 1, x.numel()).reshape(x.shape).cuda()
    hal = m(x)
    (hal * gw).sum().backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    # fp32 oracle == reference golden (context; bf16 storage noise at random init is large)
    err32 = (hal.detach().cpu() - g["hal_train"]).abs()
    print(f"\n[unet] vs fp32 reference golden: max {err32.max():.4f} mean {err32.mean():.5f}")
    assert err32.mean() < 0.03
    # bf16-storage oracle
    params = {k: v.requires_grad_(True) for k, v in state.items() if ou.is_param(k)}
    hal_e = ou.unet_forward(state, x, training=True, update_stats=True, q=ou.round_bf16)
    (hal_e * gw).sum().backward()
    err = (hal.detach() - hal_e.detach()).abs()
    print(f"[unet] vs bf16-storage oracle: hal max {err.max():.5f} mean {err.mean():.6f}")
    assert err.max().item() <= 2e-2
    keys = list(params.keys())
    c_all = cos(flat(grads, keys), flat({k: params[k].grad for k in keys}, keys))
    worst = min((cos(grads[k], params[k].grad), k) for k in keys if params[k].grad.numel() >= 4096)
    print(f"[unet] grad cosine vs bf16-storage oracle: all {c_all:.6f} worst tensor {worst}")
    assert c_all >= 0.999
    assert worst[0] >= 0.99
    for k in ("encoder.bn1.running_mean", "encoder.bn1.running_var", "decoder.blocks.4.conv2.1.running_var",
              "encoder.layer4.2.bn2.running_mean"):
        assert torch.allclose(m.state_dict()[k], state[k], rtol=2e-2, atol=2e-3), k
    assert int(m.state_dict()["encoder.bn1.num_batches_tracked"]) == 1


def test_unet_eval_forward_and_graph():
    from oracle import unet as ou
    m = make_unet()
    # non-trivial running statistics
    gen = torch.Generator().manual_seed(5)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=gen) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=gen) + 0.5)
    m.eval()
    state = oracle_state(m)
    x = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(3)).cuda()
    with torch.no_grad():
        hal = m(x)
        hal_e = ou.unet_forward(state, x, training=False, q=ou.round_bf16)
        hal_32 = ou.unet_forward(state, x, training=False)
    err = (hal - hal_e).abs()
    print(f"\n[unet eval] vs bf16-storage oracle max {err.max():.5f}; vs fp32 oracle max {(hal - hal_32).abs().max():.5f}")
    assert err.max().item() <= 2e-2
    m.use_cuda_graph = True
    with torch.no_grad():
        outs = [m(x).clone() for _ in range(3)]
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[2]) and torch.allclose(outs[0], hal, atol=1e-6)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 60, 64).cuda())
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))


def test_unet_train_cuda_graph_matches_eager():
    m = make_unet().train()
    m2 = copy.deepcopy(m)
    m2.use_cuda_graph = True
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3)).cuda()
    for step in range(3):
        for mod in (m, m2):
            mod.zero_grad(set_to_none=True)
            out = mod(x)
            out.square().sum().backward()
    torch.cuda.synchronize()
    g1 = torch.cat([p.grad.flatten() for p in m.parameters()])
    g2 = torch.cat([p.grad.flatten() for p in m2.parameters()])
    assert cos(g1, g2) > 0.9999
    assert torch.allclose(m.state_dict()["encoder.bn1.running_var"], m2.state_dict()["encoder.bn1.running_var"], rtol=1e-3)


@pytest.mark.parametrize("name", ["fasterrcnn", "retinanet"])
def test_frozen_backbone_forward_and_dgrad(name):
    from oracle import detector as odet, backbone as obb, unet as ou
    from hallucidet_b200.backbone import FrozenBackbone
    det = odet.build_detector(name, seed=123)
    odet.randomize_bn_stats(det, seed=7)
    det = det.cuda()
    state = {k: v.detach().clone() for k, v in det.backbone.state_dict().items()}
    fb = FrozenBackbone.from_torchvision(copy.deepcopy(det.backbone)).cuda()
    assert fb.out_channels == 256 and not any(p.requires_grad for p in fb.parameters())
    assert list(fb.state_dict().keys()) == list(det.backbone.state_dict().keys())
    x = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(11)).cuda().requires_grad_(True)
    feats = fb(x)
    xe = x.detach().clone().requires_grad_(True)
    feats_e = obb.backbone_forward(state, xe, variant=name, q=ou.round_bf16)
    with torch.no_grad():
        feats_32 = obb.backbone_forward(state, x.detach(), variant=name)
    assert list(feats.keys()) == list(feats_e.keys())
    gen = torch.Generator().manual_seed(2)
    loss = loss_e = 0.0
    for k in feats:
        assert feats[k].shape == feats_e[k].shape
        scale = feats_e[k].detach().abs().max().item()
        err = (feats[k].detach() - feats_e[k].detach()).abs().max().item()
        err32 = (feats[k].detach() - feats_32[k]).abs().max().item()
        print(f"\n[backbone {name}] level {k}: max|ref| {scale:.3f} err vs bf16-oracle {err:.4f} vs fp32 {err32:.4f}")
        assert err <= 2e-2 * scale
        w = torch.randn(feats[k].shape, generator=gen).cuda()
        loss = loss + (feats[k] * w).sum()
        loss_e = loss_e + (feats_e[k] * w).sum()
    loss.backward()
    loss_e.backward()
    c = cos(x.grad, xe.grad)
    print(f"[backbone {name}] d(loss)/d(image) cosine vs bf16-storage oracle {c:.6f}")
    assert c >= 0.999
    # a second, gradient-free forward (the reference's extra RGB / IR passes) must not disturb a pending backward
    x2 = torch.rand(2, 3, 128, 128).cuda().requires_grad_(True)
    f2 = fb(x2)
    with torch.no_grad():
        fb(torch.rand(2, 3, 128, 128).cuda())
    sum(v.sum() for v in f2.values()).backward()
    assert x2.grad is not None and torch.isfinite(x2.grad).all()


def test_transform_module(golden_dir):
    from hallucidet_b200.transform import GeneralizedRCNNTransform
    from oracle import step as ostep
    g = torch.load(os.path.join(golden_dir, "transform.pt"), weights_only=False)
    t = GeneralizedRCNNTransform(min_size=128, max_size=128, image_mean=[0.0], image_std=[1.0], size_divisible=1, fixed_size=(128, 128))
    _, _, targets = ostep.synthetic_batch(2, 64, 96, seed=123, device="cuda")
    imgs = g["imgs"].cuda().requires_grad_(True)
    il, tg = t(imgs, targets)
    assert torch.equal(il.tensors.detach().cpu(), g["out"])
    assert [tuple(s) for s in il.image_sizes] == [tuple(s) for s in g["image_sizes"]]
    for a, b in zip(tg, g["boxes"]):
        assert torch.equal(a["boxes"].cpu(), b)
    w = torch.randn_like(il.tensors)
    (il.tensors * w).sum().backward()
    ref = imgs.detach().clone().requires_grad_(True)
    (torch.nn.functional.interpolate(ref, size=[128, 128]) * w).sum().backward()
    assert torch.allclose(imgs.grad, ref.grad, rtol=1e-5, atol=1e-6)
    # list input with different shapes goes through the per-image path
    il2, _ = t([imgs.detach()[0], torch.rand(3, 32, 48).cuda()], None)
    assert il2.tensors.shape == (2, 3, 128, 128)


@pytest.mark.parametrize("name", ["fasterrcnn", "retinanet"])
def test_train_step_vs_oracle(name):
    from oracle import unet as ou, detector as odet, step as ostep, backbone as obb
    from hallucidet_b200.train import HalluciDetTrainer
    ir, rgb, targets = ostep.synthetic_batch(2, 64, 96, seed=123, device="cuda")
    det_cpu = odet.build_detector(name, seed=123)
    odet.randomize_bn_stats(det_cpu, seed=7)
    tr = HalluciDetTrainer(detector_name=name, size=128, pixel="mse", weights={"pixel_rgb": 1.0, "pixel_ir": 0.5}, seed=123,
                           detector_state=det_cpu.state_dict())
    state = oracle_state(tr.encoder_decoder)
    tr.encoder_decoder.train()
    out = tr.forward_step(rgb, targets, ir, targets, det_seed=7)
    out["total"].backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in tr.encoder_decoder.named_parameters()}
    det = det_cpu.cuda()
    w = {"pixel_rgb": 1.0, "pixel_ir": 0.5}
    # fp32 oracle: the loss gate (north star: within 1 % relative)
    r32 = ostep.train_step({k: v.clone() for k, v in state.items()}, det, ir, rgb, targets, size=128, detector_name=name,
                           pixel="mse", weights=w, det_seed=7)
    rel = abs(float(out["total"]) - float(r32["loss"])) / abs(float(r32["loss"]))
    print(f"\n[step {name}] loss {float(out['total']):.6f} fp32 oracle {float(r32['loss']):.6f} rel diff {rel:.5f}")
    assert rel <= 1e-2
    # bf16-storage oracle: gradients
    bstate = {k: v.detach().clone() for k, v in det.backbone.state_dict().items()}
    st_e = {k: v.clone() for k, v in state.items()}
    re = ostep.train_step(st_e, det, ir, rgb, targets, size=128, detector_name=name, pixel="mse", weights=w, det_seed=7,
                          unet_fn=lambda x: ou.unet_forward(st_e, x, training=True, q=ou.round_bf16),
                          backbone_fn=lambda x: obb.backbone_forward(bstate, x, variant=name, q=ou.round_bf16))
    keys = list(re["grads"].keys())
    c_all = cos(flat(grads, keys), flat(re["grads"], keys))
    c32 = cos(flat(grads, keys), flat(r32["grads"], keys))
    herr = (out["hal"].detach() - re["hal"]).abs().max().item()
    print(f"[step {name}] grad cosine vs bf16-storage oracle {c_all:.6f} (vs fp32 oracle {c32:.4f}); hal max err {herr:.5f}; "
          f"loss rel diff vs bf16 oracle {abs(float(out['total']) - float(re['loss'])) / abs(float(re['loss'])):.6f}")
    assert herr <= 2e-2
    assert c_all >= 0.99


def test_trainer_steps_and_reference_extra_passes():
    from oracle import step as ostep
    from hallucidet_b200.train import HalluciDetTrainer
    ir, rgb, targets = ostep.synthetic_batch(2, 64, 64, seed=1, device="cuda")
    tr = HalluciDetTrainer(detector_name="fasterrcnn", size=128, seed=123, reference_extra_passes=True)
    before = tr.encoder_decoder.segmentation_head[0].weight.detach().clone()
    losses = []
    for _ in range(3):
        out = tr.training_step(rgb, targets, ir, targets, det_seed=3)
        losses.append(float(out["total"]))
    torch.cuda.synchronize()
    assert all(l == l and l > 0 for l in losses)
    assert not torch.equal(before, tr.encoder_decoder.segmentation_head[0].weight.detach())
    assert max(float(p.grad.abs().max()) for p in tr.encoder_decoder.parameters()) <= 0.5 + 1e-6
