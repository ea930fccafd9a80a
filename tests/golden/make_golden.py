"""Generate tests/golden/*.pt by running THE REFERENCE ITSELF (read-only at /root/reference).

Run in the build container only (``python tests/golden/make_golden.py``); the GPU box has no
/root/reference and only replays the committed fixtures.  No reference source is copied: the
reference's own files are imported in-process with the three stubs of SURVEY.md section 8(c):
  (1) fake ``pretrainedmodels`` exposing ``pretrained_settings`` (encoders/resnet.py:32 imports it);
  (2) empty package shells for ``segmentation_models{,.encoders,.decoders}`` so that the real
      ``encoders/{_utils,_base,resnet}.py``, ``base/*`` and ``decoders/unet/*`` load without pulling
      timm / efficientnet (encoders/__init__.py:1);
  (3) ``get_encoder`` restricted to the resnet registry (restating encoders/__init__.py:59-85).
Detector side: ``src.models.detector.Detector`` is imported unmodified, with
``Detector.select_detector`` monkey-patched to build torchvision detectors with weights=None.

Fixtures written (all fp32, CPU, deterministic seeds):
  unet_small.pt       B=2 64x96: init checksums, hal (train & eval BN), running-stat updates, grads of a loss
  transform.pt        nearest index maps 512->640, 1024->300, 1280->300, 64->128, 96->128 + a resized image + boxes
  frcnn_small.pt      Faster R-CNN eval-forward losses + backbone features + d(loss)/d(image) at S=128
  retina_small.pt     RetinaNet ditto
  step_small.pt       the assembled forward_step loss / dhal / U-Net grads (B=2, 64x96, S=128, mse regulariser on)
"""
import importlib
import importlib.util
import os
import sys
import types

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)


def _load(name, path, package=False):
    spec = importlib.util.spec_from_file_location(
        name, path, submodule_search_locations=[os.path.dirname(path)] if package else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference_unet():
    names = ["resnet18", "resnet34", "resnet50", "resnet101", "resnet152", "resnext50_32x4d", "resnext101_32x4d",
             "resnext101_32x8d", "resnext101_32x16d", "resnext101_32x32d", "resnext101_32x48d"]
    pm = types.ModuleType("pretrainedmodels")
    pmm = types.ModuleType("pretrainedmodels.models")
    pmt = types.ModuleType("pretrainedmodels.models.torchvision_models")
    pmt.pretrained_settings = {n: {"imagenet": {"url": None}} for n in names}
    pm.models, pmm.torchvision_models = pmm, pmt
    sys.modules.update({"pretrainedmodels": pm, "pretrainedmodels.models": pmm,
                        "pretrainedmodels.models.torchvision_models": pmt})
    root = os.path.join(REF, "src", "segmentation_models")
    for pkg, sub in (("segmentation_models", ""), ("segmentation_models.encoders", "encoders"),
                     ("segmentation_models.decoders", "decoders")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(root, sub)]
        sys.modules[pkg] = m
    _load("segmentation_models.encoders._utils", os.path.join(root, "encoders", "_utils.py"))
    _load("segmentation_models.encoders._base", os.path.join(root, "encoders", "_base.py"))
    rn = _load("segmentation_models.encoders.resnet", os.path.join(root, "encoders", "resnet.py"))

    def get_encoder(name, in_channels=3, depth=5, weights=None, output_stride=32, **kwargs):
        Encoder = rn.resnet_encoders[name]["encoder"]
        params = rn.resnet_encoders[name]["params"]
        params.update(depth=depth)
        encoder = Encoder(**params)
        assert weights is None, "no network: random init only"
        encoder.set_in_channels(in_channels, pretrained=weights is not None)
        return encoder

    sys.modules["segmentation_models.encoders"].get_encoder = get_encoder
    _load("segmentation_models.base", os.path.join(root, "base", "__init__.py"), package=True)
    unet = _load("segmentation_models.decoders.unet", os.path.join(root, "decoders", "unet", "__init__.py"), package=True)
    return unet.Unet


def build_reference_unet(Unet, seed):
    """src/models/encoder_decoder.py:22-30 with encoder_weights=None."""
    torch.manual_seed(seed)
    m = Unet("resnet34", encoder_depth=5, encoder_weights=None, decoder_attention_type=None, in_channels=3, classes=3)
    m.segmentation_head[-1] = torch.nn.Sigmoid()
    return m


def import_reference_detector():
    import torchvision
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        det_mod = importlib.import_module("src.models.detector")
    finally:
        os.chdir(cwd)

    def select_detector(detector_name="fasterrcnn_resnet50_fpn", pretrained=True):
        if "retinanet" in detector_name:
            return torchvision.models.detection.retinanet_resnet50_fpn(weights=None, weights_backbone=None)
        return torchvision.models.detection.fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None)

    det_mod.Detector.select_detector = staticmethod(select_detector)
    return det_mod.Detector


def fingerprint(t):
    t = t.detach().double().flatten()
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return torch.stack([t.sum(), t.abs().sum(), (t * torch.cos(idx * 0.37)).sum()]).float()


def main():
    from oracle import unet as ounet, detector as odet, step as ostep, transform as otr

    torch.set_num_threads(8)
    Unet = import_reference_unet()
    Detector = import_reference_detector()

    # ---- U-Net ---------------------------------------------------------------------------------
    ref = build_reference_unet(Unet, 123)
    sd = ref.state_dict()
    mine = ounet.init_unet_state(123)
    assert list(sd.keys()) == list(mine.keys()), "state-dict key order differs from the reference"
    for k in sd:
        assert torch.equal(sd[k], mine[k]), f"init mismatch at {k}"
    print("oracle init == reference init (bit-exact), keys:", len(sd))
    init_fp = {k: fingerprint(v) for k, v in sd.items() if v.dtype.is_floating_point}

    ir, rgb, targets = ostep.synthetic_batch(2, 64, 96, seed=123)
    x = ir.repeat(1, 3, 1, 1)
    ref.train()
    hal_train = ref(x)
    gw = torch.linspace(-1, 1, hal_train.numel()).reshape(hal_train.shape)
    (hal_train * gw).sum().backward()
    grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
    sd_after = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    ref.eval()
    with torch.no_grad():
        hal_eval = ref(x)
    torch.save({
        "seed": 123, "ir": ir, "rgb": rgb,
        "init_fingerprint": init_fp,
        "hal_train": hal_train.detach(), "hal_eval_after_one_train_step": hal_eval,
        "grad_weight": "linspace(-1,1,numel).reshape(hal.shape)",
        "grad_fingerprint": {k: fingerprint(v) for k, v in grads.items()},
        "grad_full": {k: grads[k] for k in ("segmentation_head.0.weight", "segmentation_head.0.bias",
                                             "decoder.blocks.4.conv2.0.weight", "decoder.blocks.0.conv1.1.weight",
                                             "encoder.bn1.weight", "encoder.layer4.2.bn2.bias")},
        "grad_conv1_slice": grads["encoder.conv1.weight"][:4].clone(),
        "running_after": {k: v for k, v in sd_after.items() if "running" in k and ("bn1" in k or "blocks.4" in k)},
    }, os.path.join(HERE, "unet_small.pt"))
    print("unet_small.pt written; hal range", float(hal_train.detach().min()), float(hal_train.detach().max()))

    # ---- transform -----------------------------------------------------------------------------
    tmod = importlib.import_module("src.models.custom_generalized_transform")
    T = tmod.CustomGeneralizedRCNNTransform(min_size=128, max_size=128, image_mean=[0.0], image_std=[1.0],
                                            size_divisible=1, fixed_size=(128, 128))
    imgs = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(5))
    il, tg = T(imgs, [dict(t) for t in targets])
    maps = {}
    for o, i in ((640, 512), (640, 640), (300, 1024), (300, 1280), (128, 64), (128, 96)):
        a = torch.arange(i, dtype=torch.float32).reshape(1, 1, 1, i)
        maps[f"{i}->{o}"] = torch.nn.functional.interpolate(a, size=[1, o])[0, 0, 0].to(torch.int64)
    torch.save({"imgs": imgs, "out": il.tensors, "image_sizes": il.image_sizes, "boxes": [t["boxes"] for t in tg],
                "index_maps": maps}, os.path.join(HERE, "transform.pt"))
    print("transform.pt written")

    # ---- detectors -----------------------------------------------------------------------------
    for name, fname in (("fasterrcnn", "frcnn_small.pt"), ("retinanet", "retina_small.pt")):
        torch.manual_seed(123)
        D = Detector(name=name, pretrained=False, n_classes=2, size=128)
        det = D.detector
        mine_det = odet.build_detector(name, seed=123)
        for (k1, v1), (k2, v2) in zip(det.state_dict().items(), mine_det.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2), f"detector init mismatch {k1}"
        odet.randomize_bn_stats(det, seed=7)
        for p in det.parameters():
            p.requires_grad_(False)
        img = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(11)).requires_grad_(True)
        torch.manual_seed(7)
        losses, detections = Detector.calculate_loss(det, img, [dict(t) for t in targets], train_det=False, model_name=name)
        total = sum(v for v in losses.values())
        total.backward()
        with torch.no_grad():
            feats = det.backbone(det.transform(img.detach())[0].tensors)
        torch.save({"img": img.detach(), "losses": {k: v.detach() for k, v in losses.items()},
                    "dimg": img.grad.detach().clone(),
                    "features_fingerprint": {k: fingerprint(v) for k, v in feats.items()},
                    "feature_last": list(feats.values())[-1].detach().clone(),
                    "n_detections": [len(d["boxes"]) for d in detections]}, os.path.join(HERE, fname))
        print(fname, {k: float(v) for k, v in losses.items()})

    # ---- assembled step (train_hallucidet.py:161-209) -------------------------------------------
    torch.manual_seed(123)
    D = Detector(name="fasterrcnn", pretrained=False, n_classes=2, size=128)
    det = D.detector
    odet.randomize_bn_stats(det, seed=7)
    for p in det.parameters():
        p.requires_grad_(False)
    ref = build_reference_unet(Unet, 123)
    ref.train()
    losses_mod = importlib.import_module("src.losses.losses")
    loss_pixel = losses_mod.Reconstruction.select_loss_pixel("mse")
    w = {"pixel_rgb": 1.0, "pixel_ir": 0.5}
    ir3 = ir.repeat(1, 3, 1, 1)
    hal = ref(ir3)
    hal.retain_grad()
    l_rgb = loss_pixel(rgb, hal) * w["pixel_rgb"]
    l_ir = loss_pixel(ir3, hal) * w["pixel_ir"]
    torch.manual_seed(7)
    losses_det, _ = Detector.calculate_loss(det, hal, [dict(t) for t in targets], train_det=False, model_name="fasterrcnn")
    losses_det["classification"] = losses_det["loss_classifier"]
    losses_det["bbox_regression"] = losses_det["loss_box_reg"]
    det_total = (losses_det["bbox_regression"] * 0.1 + losses_det["classification"] * 0.1 +
                 losses_det["loss_objectness"] * 0.1 + losses_det["loss_rpn_box_reg"] * 0.1 + 0.0)
    total = det_total + l_rgb + 0.0 + l_ir + 0.0
    total.backward()
    torch.save({"loss": total.detach(), "det_total": det_total.detach(), "pixel_rgb": l_rgb.detach(), "pixel_ir": l_ir.detach(),
                "losses_det": {k: v.detach() for k, v in losses_det.items()}, "hal": hal.detach(),
                "dhal": hal.grad.detach().clone(),
                "grad_fingerprint": {k: fingerprint(p.grad) for k, p in ref.named_parameters()},
                "weights": w, "size": 128}, os.path.join(HERE, "step_small.pt"))
    print("step_small.pt loss", float(total))


if __name__ == "__main__":
    main()
