"""Pin the oracle (oracle/*.py) to outputs of the reference itself (tests/golden/*.pt, made by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import os

import pytest
import torch

from oracle import unet as ounet, detector as odet, step as ostep, transform as otr, backbone as obb


def fingerprint(t):
    t = t.detach().double().flatten()
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return torch.stack([t.sum(), t.abs().sum(), (t * torch.cos(idx * 0.37)).sum()]).float()


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


@pytest.fixture(scope="module")
def unet_state():
    torch.set_num_threads(8)
    return ounet.init_unet_state(123)


def test_unet_init_matches_reference(golden_dir, unet_state):
    g = load(golden_dir, "unet_small.pt")
    assert len(unet_state) == 278
    assert sum(v.numel() for k, v in unet_state.items() if ounet.is_param(k)) == 24436659
    for k, fp in g["init_fingerprint"].items():
        assert torch.equal(fingerprint(unet_state[k]), fp), k


def test_unet_forward_backward_matches_reference(golden_dir):
    g = load(golden_dir, "unet_small.pt")
    state = ounet.init_unet_state(123)
    params = {k: v.requires_grad_(True) for k, v in state.items() if ounet.is_param(k)}
    x = ounet.expand_ir(g["ir"], 3)
    hal = ounet.unet_forward(state, x, training=True, update_stats=True)
    assert torch.allclose(hal, g["hal_train"], atol=1e-6, rtol=0)
    gw = torch.linspace(-1, 1, hal.numel()).reshape(hal.shape)
    (hal * gw).sum().backward()
    for k, full in g["grad_full"].items():
        assert torch.allclose(params[k].grad, full, atol=1e-5, rtol=1e-4), k
    assert torch.allclose(params["encoder.conv1.weight"].grad[:4], g["grad_conv1_slice"], atol=1e-5, rtol=1e-4)
    for k, fp in g["grad_fingerprint"].items():
        mine = fingerprint(params[k].grad)
        assert torch.allclose(mine, fp, rtol=2e-3, atol=2e-4), (k, mine, fp)
    for k, v in g["running_after"].items():
        assert torch.allclose(state[k], v, atol=1e-6), k
    with torch.no_grad():
        hal_eval = ounet.unet_forward(state, x, training=False)
    assert torch.allclose(hal_eval, g["hal_eval_after_one_train_step"], atol=1e-6, rtol=0)


def test_unet_rejects_bad_shape(unet_state):
    with pytest.raises(RuntimeError):
        ounet.unet_forward(unet_state, torch.zeros(1, 3, 60, 64), training=False)


def test_transform_matches_reference(golden_dir):
    g = load(golden_dir, "transform.pt")
    for key, ref_map in g["index_maps"].items():
        i, o = (int(s) for s in key.split("->"))
        assert torch.equal(otr.nearest_src_index(o, i), ref_map), key
    _, _, targets = ostep.synthetic_batch(2, 64, 96, seed=123)
    out, sizes, tg = otr.transform_forward(g["imgs"], targets, size=128)
    assert torch.equal(out, g["out"])
    assert [tuple(s) for s in sizes] == [tuple(s) for s in g["image_sizes"]]
    for a, b in zip(tg, g["boxes"]):
        assert torch.equal(a["boxes"], b)


@pytest.mark.parametrize("name,fname", [("fasterrcnn", "frcnn_small.pt"), ("retinanet", "retina_small.pt")])
def test_detector_loss_and_dgrad_match_reference(golden_dir, name, fname):
    g = load(golden_dir, fname)
    _, _, targets = ostep.synthetic_batch(2, 64, 96, seed=123)
    det = odet.build_detector(name, seed=123)
    odet.randomize_bn_stats(det, seed=7)
    img = g["img"].clone().requires_grad_(True)
    torch.manual_seed(7)
    losses, detections = odet.calculate_loss(det, img, targets, 128, name)
    for k, v in g["losses"].items():
        assert torch.allclose(losses[k], v, rtol=1e-5, atol=1e-6), (k, losses[k], v)
    sum(losses.values()).backward()
    assert torch.allclose(img.grad, g["dimg"], rtol=1e-3, atol=1e-7)
    assert [len(d["boxes"]) for d in detections] == g["n_detections"]
    with torch.no_grad():
        batched, _, _ = otr.transform_forward(g["img"], None, size=128)
        feats = obb.backbone_forward(det.backbone.state_dict(), batched, variant=name)
    assert list(feats.keys()) == list(g["features_fingerprint"].keys())
    assert torch.allclose(list(feats.values())[-1], g["feature_last"], atol=1e-3, rtol=1e-4)
    for k, fp in g["features_fingerprint"].items():
        assert torch.allclose(fingerprint(feats[k]), fp, rtol=1e-3, atol=1e-1), k


def test_assembled_step_matches_reference(golden_dir):
    g = load(golden_dir, "step_small.pt")
    ir, rgb, targets = ostep.synthetic_batch(2, 64, 96, seed=123)
    state = ounet.init_unet_state(123)
    det = odet.build_detector("fasterrcnn", seed=123)
    odet.randomize_bn_stats(det, seed=7)
    r = ostep.train_step(state, det, ir, rgb, targets, size=g["size"], pixel="mse", weights=g["weights"], det_seed=7)
    assert torch.allclose(r["hal"], g["hal"], atol=1e-6)
    assert torch.allclose(r["loss"], g["loss"], rtol=1e-5)
    assert torch.allclose(r["pixel_rgb"].detach(), g["pixel_rgb"], rtol=1e-6)
    assert torch.allclose(r["pixel_ir"].detach(), g["pixel_ir"], rtol=1e-6)
    assert torch.allclose(r["dhal"], g["dhal"], rtol=1e-3, atol=1e-8)
    cos_num = cos_a = cos_b = 0.0
    for k, fp in g["grad_fingerprint"].items():
        mine = fingerprint(r["grads"][k])
        assert torch.allclose(mine, fp, rtol=5e-3, atol=5e-5), (k, mine, fp)
