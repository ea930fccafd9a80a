"""Oracle: frozen torchvision ResNet-50-FPN detector backbone as functional fp32 PyTorch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``state`` has the keys of
``detector.backbone.state_dict()`` (``body.*``, ``fpn.*``).

Follows (TV: = installed torchvision, the third-party code the reference calls at
src/utils/eval_forward_fasterrcnn.py:55 and src/utils/eval_forward_retinanet.py:125):
  TV: models/detection/backbone_utils.py:56-59     BackboneWithFPN.forward
  TV: models/resnet.py:108-163                     Bottleneck (stride on conv2)
  TV: ops/misc.py:14-61                            FrozenBatchNorm2d  (== eval-mode BatchNorm2d)
  TV: ops/feature_pyramid_network.py:172-250       FPN forward, LastLevelMaxPool, LastLevelP6P7
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

LAYERS = (3, 4, 6, 3)
PLANES = (64, 128, 256, 512)


def _id(x):
    return x


def _conv_bn(state, conv, bn, x, eps, q, stride=1, padding=0):
    """conv followed by frozen BN.  With a rounding hook q the BN scale is folded into the weights BEFORE rounding
    (what a frozen-weight kernel stores); mathematically identical to conv -> affine."""
    w_, b_ = state[bn + ".weight"], state[bn + ".bias"]
    rm, rv = state[bn + ".running_mean"], state[bn + ".running_var"]
    scale = w_ * (rv + eps).rsqrt()
    shift = b_ - rm * scale
    if q is _id:
        return _frozen_bn(state, bn, F.conv2d(x, state[conv + ".weight"], stride=stride, padding=padding), eps)
    w = q(state[conv + ".weight"] * scale.reshape(-1, 1, 1, 1))
    return F.conv2d(x, w, stride=stride, padding=padding) + shift.reshape(1, -1, 1, 1)


def _frozen_bn(state, p, x, eps):
    """TV ops/misc.py:54-63 -- per-channel affine with the stored statistics."""
    w, b = state[p + ".weight"], state[p + ".bias"]
    rm, rv = state[p + ".running_mean"], state[p + ".running_var"]
    scale = w * (rv + eps).rsqrt()
    shift = b - rm * scale
    return x * scale.reshape(1, -1, 1, 1) + shift.reshape(1, -1, 1, 1)


def _bottleneck(state, p, x, stride, has_down, eps, q=_id):
    out = q(F.relu(_conv_bn(state, p + ".conv1", p + ".bn1", x, eps, q)))
    out = q(F.relu(_conv_bn(state, p + ".conv2", p + ".bn2", out, eps, q, stride=stride, padding=1)))
    out = _conv_bn(state, p + ".conv3", p + ".bn3", out, eps, q)
    if has_down:
        idn = q(_conv_bn(state, p + ".downsample.0", p + ".downsample.1", x, eps, q, stride=stride))
    else:
        idn = x
    return q(F.relu(out + idn))


def body_forward(state, x, eps=1e-5, q=_id):
    """ResNet-50 body -> [C2, C3, C4, C5]."""
    h = q(F.relu(_conv_bn(state, "body.conv1", "body.bn1", q(x), eps, q, stride=2, padding=3)))
    h = F.max_pool2d(h, kernel_size=3, stride=2, padding=1)
    outs = []
    for li, nblocks in enumerate(LAYERS, start=1):
        for b in range(nblocks):
            stride = 2 if (li > 1 and b == 0) else 1
            h = _bottleneck(state, f"body.layer{li}.{b}", h, stride, b == 0, eps, q)
        outs.append(h)
    return outs


def backbone_forward(state, x, variant="fasterrcnn", eps=1e-5, return_body=False, q=_id):
    """x: [B,3,S,S] fp32 -> OrderedDict of 256-channel maps.

    fasterrcnn: keys '0','1','2','3','pool' (returned layers 1-4 + LastLevelMaxPool).
    retinanet : keys '0','1','2','p6','p7'  (returned layers 2-4 + LastLevelP6P7(256,256)).
    """
    c = body_forward(state, x, eps, q)
    feats = c if variant == "fasterrcnn" else c[1:]
    n = len(feats)
    last_inner = q(F.conv2d(feats[-1], q(state[f"fpn.inner_blocks.{n-1}.0.weight"]), state[f"fpn.inner_blocks.{n-1}.0.bias"]))
    results = [F.conv2d(last_inner, q(state[f"fpn.layer_blocks.{n-1}.0.weight"]), state[f"fpn.layer_blocks.{n-1}.0.bias"], padding=1)]
    for idx in range(n - 2, -1, -1):
        lateral = q(F.conv2d(feats[idx], q(state[f"fpn.inner_blocks.{idx}.0.weight"]), state[f"fpn.inner_blocks.{idx}.0.bias"]))
        top_down = F.interpolate(last_inner, size=lateral.shape[-2:], mode="nearest")
        last_inner = q(lateral + top_down)
        results.insert(0, F.conv2d(last_inner, q(state[f"fpn.layer_blocks.{idx}.0.weight"]), state[f"fpn.layer_blocks.{idx}.0.bias"], padding=1))
    names = [str(i) for i in range(n)]
    if variant == "fasterrcnn":
        results.append(F.max_pool2d(results[-1], kernel_size=1, stride=2, padding=0))
        names.append("pool")
    else:
        p6 = F.conv2d(q(results[-1]), q(state["fpn.extra_blocks.p6.weight"]), state["fpn.extra_blocks.p6.bias"], stride=2, padding=1)
        p7 = F.conv2d(q(F.relu(q(p6))), q(state["fpn.extra_blocks.p7.weight"]), state["fpn.extra_blocks.p7.bias"], stride=2, padding=1)
        results += [p6, p7]
        names += ["p6", "p7"]
    out = OrderedDict(zip(names, results))
    if return_body:
        return out, c
    return out
