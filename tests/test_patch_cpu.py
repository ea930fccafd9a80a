"""The shipped binding (hallucidet_b200/patch.py): with the reference checkout on sys.path and the
patch applied, the REFERENCE's own ``EncoderDecoder`` / ``Detector`` classes (src/models/encoder_decoder.py:8-30,
src/models/detector.py:23-79) construct the B200 modules, and reference-shaped state dicts load into them.

Needs /root/reference (build container only; skipped on the GPU box).  Nothing is computed: construction and
``load_state_dict`` run on the CPU, the kernels are not called.
"""
import os
import subprocess
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r'''
import sys, types, json
sys.path.insert(0, {root!r}); sys.path.insert(0, {ref!r})
import torch, torchvision
from hallucidet_b200 import patch
from hallucidet_b200.unet import Unet as B200Unet
from hallucidet_b200.backbone import FrozenBackbone
from hallucidet_b200.transform import CustomGeneralizedRCNNTransform as B200Transform
bound = patch.apply()
import src.models.encoder_decoder as ed                         # the reference's own file
import src.models.detector as rd
out = {{"bound": bound}}
# --- U-Net through the reference's EncoderDecoder (encoder_decoder.py:22-30)
torch.manual_seed(123)
m = ed.EncoderDecoder(name="resnet34", encoder_depth=5, encoder_weights=None, decoder_attention_type=None, in_channels=3,
                      output_channels=3, segmentation_head="sigmoid").encoder_decoder
out["unet_is_b200"] = isinstance(m, B200Unet)
out["head_is_sigmoid"] = isinstance(m.segmentation_head[-1], torch.nn.Sigmoid)
from oracle import unet as ou                                    # reference-shaped state dict (pinned to the reference goldens)
state = ou.init_unet_state(7)
missing, unexpected = m.load_state_dict(state, strict=True)
out["unet_keys"] = len(m.state_dict())
out["unet_loaded"] = bool(torch.equal(m.state_dict()["decoder.blocks.0.conv1.0.weight"], state["decoder.blocks.0.conv1.0.weight"]))
# --- detector through the reference's Detector (detector.py:23-79); offline builder instead of the downloading one
def offline(detector_name="fasterrcnn_resnet50_fpn", pretrained=True):
    f = torchvision.models.detection.retinanet_resnet50_fpn if "retinanet" in detector_name else torchvision.models.detection.fasterrcnn_resnet50_fpn
    return f(weights=None, weights_backbone=None)
rd.Detector.select_detector = staticmethod(offline)
for name in ("fasterrcnn", "retinanet"):
    tv = offline(name)
    det = rd.Detector(name=name, pretrained=False, n_classes=2, size=128).detector
    out[name] = {{
        "backbone_is_b200": isinstance(det.backbone, FrozenBackbone),
        "transform_is_b200": isinstance(det.transform, B200Transform),
        "frozen": not any(p.requires_grad for p in det.parameters()),
        "eval": not det.training,
        "backbone_keys_equal": list(det.backbone.state_dict().keys()) == list(tv.backbone.state_dict().keys()),
    }}
    # a torchvision-shaped backbone checkpoint loads AFTER the swap (train_hallucidet.py:107-115 order)
    sd = {{k: torch.randn_like(v) if v.is_floating_point() else v for k, v in tv.backbone.state_dict().items()}}
    det.backbone.load_state_dict(sd, strict=True)
    out[name]["reload_ok"] = bool(torch.equal(det.backbone.state_dict()["body.layer1.0.conv1.weight"], sd["body.layer1.0.conv1.weight"]))
out["calc_loss_rebound"] = rd.eval_forward_fasterrcnn.__module__
print("RESULT " + json.dumps(out))
'''


def test_reference_classes_construct_b200_modules():
    code = _CHILD.format(root=ROOT, ref=REF)
    # a child process: the patch registers modules under the reference's package names
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REF, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    import json
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out["unet_is_b200"] and out["head_is_sigmoid"] and out["unet_loaded"] and out["unet_keys"] == 278
    for name in ("fasterrcnn", "retinanet"):
        assert all(out[name].values()), (name, out[name])
    assert out["calc_loss_rebound"] == "hallucidet_b200.detection"
    assert any(b.endswith("segmentation_models") for b in out["bound"])


def test_install_converts_existing_modules():
    """patch.install(): an already built smp-style U-Net (here: the oracle's reference-shaped state in a torchvision/smp
    layout stand-in is not available on this box, so the B200 module built from a state dict) keeps its weights."""
    from hallucidet_b200 import patch
    from hallucidet_b200.unet import Unet
    import torchvision

    class Lit:
        pass

    lit = Lit()
    torch.manual_seed(3)
    lit.encoder_decoder = Unet("resnet34", classes=3)
    lit.detector = torchvision.models.detection.fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None, num_classes=2)
    before = lit.encoder_decoder
    patch.install(lit)
    from hallucidet_b200.backbone import FrozenBackbone
    assert lit.encoder_decoder is before                       # already the B200 class: untouched
    assert isinstance(lit.detector.backbone, FrozenBackbone) and not lit.detector.training


def test_run_launcher_usage():
    res = subprocess.run([sys.executable, "-m", "hallucidet_b200.run"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert res.returncode != 0 and "hallucidet_b200.run" in (res.stderr + res.stdout)
