"""Oracle: the hallucination U-Net (smp.Unet, ResNet-34 encoder) as functional fp32 PyTorch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Operates on a flat ``state`` dict that has
exactly the reference's ``state_dict()`` keys/shapes (278 entries, 24,436,659 parameters).

Reference files followed (relative to the reference repo root):
  src/segmentation_models/base/model.py:24-38        SegmentationModel.forward
  src/segmentation_models/encoders/resnet.py:47-65   ResNetEncoder.get_stages / forward
  TV: models/resnet.py:59-105                        BasicBlock
  src/segmentation_models/decoders/unet/decoder.py:7-8,38-46,111-124
  src/segmentation_models/base/modules.py:10-47      Conv2dReLU = conv(bias=False)+BN+ReLU
  src/segmentation_models/base/heads.py:21-27        SegmentationHead
  src/models/encoder_decoder.py:29-30                head activation := Sigmoid
  src/segmentation_models/base/initialization.py:4-27
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

ENCODER_LAYERS = (3, 4, 6, 3)                      # encoders/resnet.py:136-144 ("resnet34")
ENCODER_CHANNELS = (3, 64, 64, 128, 256, 512)
DECODER_CHANNELS = (256, 128, 64, 32, 16)          # decoders/unet/model.py:62
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def decoder_block_channels():
    """(in, skip, out) per decoder block -- decoders/unet/decoder.py:86-96."""
    enc = list(ENCODER_CHANNELS[1:])[::-1]
    ins = [enc[0]] + list(DECODER_CHANNELS[:-1])
    skips = enc[1:] + [0]
    return list(zip(ins, skips, DECODER_CHANNELS))


def init_unet_state(seed=123, classes=3):
    """Random-init state dict, drawing the RNG in the order the reference constructor does.

    Reference order (decoders/unet/model.py:69-100): torchvision ``ResNet.__init__`` (every
    ``nn.Conv2d``/``nn.Linear`` constructor draws, then kaiming_normal_(fan_out) over all convs,
    TV models/resnet.py:208-216), ``del fc``; five DecoderBlocks (conv1, conv2 constructors);
    SegmentationHead conv; then ``initialize()``: kaiming_uniform_(fan_in, relu) on decoder convs
    and xavier_uniform_ on the head (initialization.py:4-27).
    """
    from torchvision.models.resnet import ResNet, BasicBlock

    torch.manual_seed(seed)
    enc = ResNet(block=BasicBlock, layers=list(ENCODER_LAYERS))
    dec_convs, dec_bns = [], []
    for cin, cskip, cout in decoder_block_channels():
        c1 = nn.Conv2d(cin + cskip, cout, 3, padding=1, bias=False)
        b1 = nn.BatchNorm2d(cout)
        c2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        b2 = nn.BatchNorm2d(cout)
        dec_convs.append((c1, c2))
        dec_bns.append((b1, b2))
    head = nn.Conv2d(DECODER_CHANNELS[-1], classes, 3, padding=1)
    for c1, c2 in dec_convs:
        nn.init.kaiming_uniform_(c1.weight, mode="fan_in", nonlinearity="relu")
        nn.init.kaiming_uniform_(c2.weight, mode="fan_in", nonlinearity="relu")
    nn.init.xavier_uniform_(head.weight)
    nn.init.constant_(head.bias, 0)

    state = OrderedDict()
    for k, v in enc.state_dict().items():
        if k.startswith("fc."):
            continue
        state["encoder." + k] = v.detach().clone()
    for i, ((c1, c2), (b1, b2)) in enumerate(zip(dec_convs, dec_bns)):
        for j, (c, b) in enumerate(((c1, b1), (c2, b2)), start=1):
            state[f"decoder.blocks.{i}.conv{j}.0.weight"] = c.weight.detach().clone()
            for kk, vv in b.state_dict().items():
                state[f"decoder.blocks.{i}.conv{j}.1.{kk}"] = vv.detach().clone()
    state["segmentation_head.0.weight"] = head.weight.detach().clone()
    state["segmentation_head.0.bias"] = head.bias.detach().clone()
    return state


class _RoundBF16(torch.autograd.Function):
    """bf16 storage emulation: round the value forward and the gradient backward (straight-through)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def round_bf16(x):
    return _RoundBF16.apply(x)


def _id(x):
    return x


def is_param(key):
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


def _bn(state, prefix, x, training, update_stats):
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    if training and not update_stats:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, state[prefix + ".weight"], state[prefix + ".bias"],
                     training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
    if training and update_stats:
        state[prefix + ".num_batches_tracked"] += 1
    return y


def _basic_block(state, p, x, stride, has_down, training, update_stats, q=_id):
    """TV models/resnet.py:89-105.  q = storage-rounding hook (identity for the fp32 oracle)."""
    out = q(F.conv2d(x, q(state[p + ".conv1.weight"]), stride=stride, padding=1))
    out = q(F.relu(_bn(state, p + ".bn1", out, training, update_stats)))
    out = q(F.conv2d(out, q(state[p + ".conv2.weight"]), padding=1))
    out = _bn(state, p + ".bn2", out, training, update_stats)
    if has_down:
        idn = q(F.conv2d(x, q(state[p + ".downsample.0.weight"]), stride=stride))
        idn = _bn(state, p + ".downsample.1", idn, training, update_stats)
    else:
        idn = x
    return q(F.relu(out + idn))


def encoder_forward(state, x, training=True, update_stats=False, q=_id):
    """encoders/resnet.py:47-65 -> list of 6 features."""
    feats = [x]
    h = q(F.conv2d(q(x), q(state["encoder.conv1.weight"]), stride=2, padding=3))
    h = q(F.relu(_bn(state, "encoder.bn1", h, training, update_stats)))
    feats.append(h)
    h = F.max_pool2d(h, kernel_size=3, stride=2, padding=1)
    for li, nblocks in enumerate(ENCODER_LAYERS, start=1):
        for b in range(nblocks):
            stride = 2 if (li > 1 and b == 0) else 1
            h = _basic_block(state, f"encoder.layer{li}.{b}", h, stride, li > 1 and b == 0, training, update_stats, q)
        feats.append(h)
    return feats


def upsample2x(x):
    """decoders/unet/decoder.py:7-8 (pixel replication)."""
    return x[:, :, :, None, :, None].expand(-1, -1, -1, 2, -1, 2).reshape(x.size(0), x.size(1), x.size(2) * 2, x.size(3) * 2)


def decoder_forward(state, feats, training=True, update_stats=False, q=_id):
    """decoders/unet/decoder.py:111-124, 38-46."""
    feats = feats[1:][::-1]
    x, skips = feats[0], feats[1:]
    for i in range(len(DECODER_CHANNELS)):
        x = upsample2x(x)
        if i < len(skips):
            x = torch.cat([x, skips[i]], dim=1)
        for j in (1, 2):
            p = f"decoder.blocks.{i}.conv{j}"
            x = q(F.conv2d(x, q(state[p + ".0.weight"]), padding=1))
            x = q(F.relu(_bn(state, p + ".1", x, training, update_stats)))
    return x


def unet_forward(state, x, training=True, update_stats=False, return_intermediates=False, q=_id):
    """base/model.py:24-38 with the Sigmoid head of src/models/encoder_decoder.py:29-30.

    x: [B,3,H,W] fp32, H%32==W%32==0 (base/model.py:12-22).  Returns hal [B,3,H,W] in (0,1).
    """
    h, w = x.shape[-2:]
    if h % 32 != 0 or w % 32 != 0:
        raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and width divisible by 32.")
    feats = encoder_forward(state, x, training, update_stats, q)
    d = decoder_forward(state, feats, training, update_stats, q)
    logits = F.conv2d(d, q(state["segmentation_head.0.weight"]), state["segmentation_head.0.bias"], padding=1)
    hal = torch.sigmoid(logits)
    if return_intermediates:
        return hal, {"features": feats, "decoder_out": d, "logits": logits}
    return hal


def expand_ir(ir, channels=3):
    """src/utils/utils.py:52-53."""
    return ir.repeat(1, channels, 1, 1)
