"""Host-side mirror of the reference's detector glue, so the hot path has its caller on a box where the
reference repository is absent (bench / smoke / tests).  Same names, arguments and behaviour as

  src/models/detector.py:25-66,104-141          Detector (2-class re-heading, fixed-size transform, calculate_loss)
  src/utils/eval_forward_fasterrcnn.py:13-147   eval_forward_fasterrcnn / rpn_eval / roi_heads_eval
  src/utils/eval_forward_retinanet.py:22-244    eval_forward_retinanet and its losses

The RPN / RoI heads / RetinaNet head, anchor generator, matcher, samplers, box coder and the losses are
torchvision's own objects ("stay as the reference implements them"); only ``model.transform`` and
``model.backbone`` are the B200 modules.
"""
import contextlib
import math
import os as _os

import numpy as np
from collections import OrderedDict

import torch
import torch.nn.functional as F
import torchvision
from torchvision.models.detection.roi_heads import fastrcnn_loss
from torchvision.models.detection.rpn import concat_box_prediction_layers

from . import heads, ops
from .backbone import FrozenBackbone
from .transform import GeneralizedRCNNTransform


def _xavier_init(module):
    for layer in module.modules():
        if isinstance(layer, torch.nn.Conv2d):
            torch.nn.init.xavier_uniform_(layer.weight)
            if layer.bias is not None:
                torch.nn.init.constant_(layer.bias, 0.0)


class Detector:
    """src/models/detector.py:23-79 with random-init weights unless a state dict is loaded afterwards."""

    def __init__(self, name="fasterrcnn_resnet50_fpn", pretrained=False, n_classes=2, size=300, eval_path=None, b200=True):
        self.detector = Detector.select_detector(detector_name=name, pretrained=pretrained)
        self.detector.transform = GeneralizedRCNNTransform(min_size=size, max_size=size, image_mean=[0.0], image_std=[1.0],
                                                           size_divisible=1, fixed_size=(size, size))
        if "fasterrcnn" in name:
            in_features = self.detector.roi_heads.box_predictor.cls_score.in_features
            self.detector.roi_heads.box_predictor = torchvision.models.detection.faster_rcnn.FastRCNNPredictor(in_features, n_classes)
            _xavier_init(self.detector.roi_heads)
        elif "retinanet" in name:
            out_channels = self.detector.head.classification_head.conv[0].out_channels
            num_anchors = self.detector.head.classification_head.num_anchors
            self.detector.head.classification_head.num_classes = n_classes
            cls_logits = torch.nn.Conv2d(out_channels, num_anchors * n_classes, kernel_size=3, stride=1, padding=1)
            torch.nn.init.normal_(cls_logits.weight, std=0.01)
            torch.nn.init.constant_(cls_logits.bias, -math.log((1 - 0.01) / 0.01))
            self.detector.head.classification_head.cls_logits = cls_logits
        if eval_path is not None:
            self.detector.load_state_dict(torch.load(eval_path))

    @staticmethod
    def select_detector(detector_name="fasterrcnn_resnet50_fpn", pretrained=False):
        if pretrained:
            raise NotImplementedError("no network: load a detector state dict instead of pretrained=True")
        if "retinanet" in detector_name:
            return torchvision.models.detection.retinanet_resnet50_fpn(weights=None, weights_backbone=None)
        if "fasterrcnn" in detector_name:
            return torchvision.models.detection.fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None)
        raise NotImplementedError(f"{detector_name}: the B200 hot path covers fasterrcnn and retinanet")

    @staticmethod
    def calculate_loss(detector, outs, targets, train_det=False, model_name="fasterrcnn"):
        if "fasterrcnn" in model_name:
            return eval_forward_fasterrcnn(detector, outs, targets, train_det=train_det, model_name=model_name)
        if "retinanet" in model_name:
            return eval_forward_retinanet(detector, outs, targets, train_det=train_det, model_name=model_name)
        raise NotImplementedError(model_name)


FUSED_HEAD_CONV_RELU = _os.environ.get("HD_FUSED_HEAD", "1") == "1"


class _ConvBiasReLU(torch.autograd.Function):
    """conv + bias + ReLU of a FROZEN head convolution as one cuDNN call (the bias add and the ReLU are otherwise two more
    passes over the largest feature maps of the step); backward: ReLU mask, then the input gradient only."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups):
        out = torch.cudnn_convolution_relu(x, weight, bias, stride, padding, dilation, groups)
        ctx.save_for_backward(x, weight, out)
        ctx.cfg = (stride, padding, dilation, groups)
        return out

    @staticmethod
    def backward(ctx, grad):
        x, weight, out = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.cfg
        grad = torch.ops.aten.threshold_backward(grad, out, 0)
        gx = torch.ops.aten.convolution_backward(grad, x, weight, None, stride, padding, dilation, False, [0, 0], groups,
                                                 [True, False, False])[0]
        return gx, None, None, None, None, None, None


class _FusedConvReLU(torch.nn.Sequential):
    """Drop-in for torchvision's ``Conv2dNormActivation(conv, ReLU)`` inside the frozen detector heads (same children, same
    state_dict keys); forward = _ConvBiasReLU when the convolution is frozen and the input is an fp32 CUDA tensor."""

    def forward(self, x):
        conv = self[0]
        if (FUSED_HEAD_CONV_RELU and x.is_cuda and x.dtype == torch.float32 and not conv.weight.requires_grad
                and (conv.bias is None or not conv.bias.requires_grad) and conv.padding_mode == "zeros"
                and not isinstance(conv.padding, str)):
            bias = conv.bias if conv.bias is not None else conv.weight.new_zeros(conv.out_channels)
            return _ConvBiasReLU.apply(x, conv.weight, bias, list(conv.stride), list(conv.padding), list(conv.dilation), conv.groups)
        return super().forward(x)


def _fuse_head_conv_relu(module):
    for name, child in list(module.named_children()):
        if (isinstance(child, torch.nn.Sequential) and not isinstance(child, _FusedConvReLU) and len(child) == 2
                and isinstance(child[0], torch.nn.Conv2d) and isinstance(child[1], torch.nn.ReLU)):
            setattr(module, name, _FusedConvReLU(child[0], child[1]))
        else:
            _fuse_head_conv_relu(child)


def install_b200_backbone(detector):
    """Freeze the detector and replace ``detector.backbone`` by the B200 dgrad-only module (call AFTER weights are loaded)."""
    detector.eval()
    for p in detector.parameters():
        p.requires_grad_(False)
    if not isinstance(detector.backbone, FrozenBackbone):
        detector.backbone = FrozenBackbone.from_torchvision(detector.backbone)
    from . import backbone as _bb
    if _bb.CHANNELS_LAST_FEATURES:
        # the backbone hands over channels-last feature maps: keep the head convolutions' weights in the same format so
        # cuDNN runs them without per-call layout conversions (values and state_dict are unchanged)
        for name in ("rpn", "head"):
            head = getattr(detector, name, None)
            if isinstance(head, torch.nn.Module):
                for m in head.modules():
                    if isinstance(m, torch.nn.Conv2d):
                        m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
    for name in ("rpn", "head"):
        head = getattr(detector, name, None)
        if isinstance(head, torch.nn.Module):
            _fuse_head_conv_relu(head)
    return detector


def _check_targets(targets):
    for target in targets:
        boxes = target["boxes"]
        torch._assert(isinstance(boxes, torch.Tensor), f"Expected target boxes to be of type Tensor, got {type(boxes)}.")
        torch._assert(len(boxes.shape) == 2 and boxes.shape[-1] == 4,
                      f"Expected target boxes to be a tensor of shape [N, 4], got {boxes.shape}.")


_SIDE_STREAMS = {}


def _side_streams(device, n):
    key = (device.type, device.index)
    pool = _SIDE_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def batched_nms(boxes, scores, idxs, iou_threshold):
    """``torchvision.ops.batched_nms`` (TV ops/boxes.py: the coordinate-offset trick, then nms) with the NMS itself on
    this package's kernels (ops.nms: identical keep set).  Oversized / non-CUDA inputs go to torchvision unchanged."""
    from torchvision.ops import boxes as box_ops
    if (not boxes.is_cuda) or boxes.dtype != torch.float32 or boxes.shape[0] > ops.NMS_MAX_BOXES or boxes.numel() > 100_000:
        return box_ops.batched_nms(boxes, scores, idxs, iou_threshold)
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + 1)
    boxes_for_nms = boxes + offsets[:, None]
    return ops.nms(boxes_for_nms, scores, iou_threshold)


def _clip_boxes_batched(boxes, image_shapes):
    """``clip_boxes_to_image`` (TV ops/boxes.py) for a batch [B, ..., 4]: scalar clamp when every image has the same
    size (what the detector transform produces), per-image bounds otherwise -- the same min(max(x, 0), size) either way."""
    dim = boxes.dim()
    bx, by = boxes[..., 0::2], boxes[..., 1::2]
    if all(tuple(s) == tuple(image_shapes[0]) for s in image_shapes):
        height, width = image_shapes[0]
        bx = bx.clamp(min=0, max=width)
        by = by.clamp(min=0, max=height)
    else:
        sizes = torch.tensor([list(s) for s in image_shapes], dtype=boxes.dtype, device=boxes.device)
        shape = [len(image_shapes)] + [1] * (dim - 1)
        h, w = sizes[:, 0].view(shape), sizes[:, 1].view(shape)
        zero = torch.zeros((), dtype=boxes.dtype, device=boxes.device)
        bx = torch.minimum(torch.maximum(bx, zero), w)
        by = torch.minimum(torch.maximum(by, zero), h)
    return torch.stack((bx, by), dim=dim).reshape(boxes.shape)


_CODER_WEIGHTS = {}


def _coder_weights(box_coder, dtype, device):
    """BoxCoder.weights as a cached device tensor (torchvision's encode_single re-creates it -- a host->device copy --
    on every call)."""
    key = (tuple(box_coder.weights), dtype, str(device))
    w = _CODER_WEIGHTS.get(key)
    if w is None:
        w = _CODER_WEIGHTS[key] = torch.as_tensor(box_coder.weights, dtype=dtype, device=device)
    return w


def _encode_single(box_coder, reference_boxes, proposals):
    from torchvision.models.detection._utils import encode_boxes
    return encode_boxes(reference_boxes, proposals, _coder_weights(box_coder, reference_boxes.dtype, reference_boxes.device))


def _decode(box_coder, rel_codes, boxes):
    """``BoxCoder.decode`` + ``decode_single`` (TV models/detection/_utils.py:162-226) with the two 0.5 factors as Python
    scalars: torchvision builds ``torch.tensor(0.5, device=...)`` twice per call, i.e. two blocking host->device copies.
    0.5 is exact in every float type, so the products are identical."""
    boxes_per_image = [b.size(0) for b in boxes]
    concat_boxes = torch.cat(boxes, dim=0)
    box_sum = sum(boxes_per_image)
    if box_sum > 0:
        rel_codes = rel_codes.reshape(box_sum, -1)
    b = concat_boxes.to(rel_codes.dtype)
    widths = b[:, 2] - b[:, 0]
    heights = b[:, 3] - b[:, 1]
    ctr_x = b[:, 0] + 0.5 * widths
    ctr_y = b[:, 1] + 0.5 * heights
    wx, wy, ww, wh = box_coder.weights
    dx = rel_codes[:, 0::4] / wx
    dy = rel_codes[:, 1::4] / wy
    dw = rel_codes[:, 2::4] / ww
    dh = rel_codes[:, 3::4] / wh
    dw = torch.clamp(dw, max=box_coder.bbox_xform_clip)
    dh = torch.clamp(dh, max=box_coder.bbox_xform_clip)
    pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
    pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
    pred_w = torch.exp(dw) * widths[:, None]
    pred_h = torch.exp(dh) * heights[:, None]
    c_to_c_h = 0.5 * pred_h
    c_to_c_w = 0.5 * pred_w
    pred_boxes = torch.stack((pred_ctr_x - c_to_c_w, pred_ctr_y - c_to_c_h, pred_ctr_x + c_to_c_w, pred_ctr_y + c_to_c_h), dim=2).flatten(1)
    if box_sum > 0:
        pred_boxes = pred_boxes.reshape(box_sum, -1, 4)
    return pred_boxes


_DEFERRED_CHECKS = []


def _run_deferred_checks():
    """Asynchronous input checks (flag copied to pinned memory when the check was issued) are evaluated at the next host
    sync the step performs anyway."""
    while _DEFERRED_CHECKS:
        event, flag, on_fail = _DEFERRED_CHECKS.pop(0)
        event.synchronize()
        if bool(flag.item()):
            on_fail()


class _Pending:
    """Device work has been enqueued; the host still needs ``counts_dev`` (data-dependent sizes) to finish.  Several
    pending results are resolved with ONE device->host read (``_resolve``), so independent parts of the tail share a sync."""

    def __init__(self, counts_dev, finish):
        self.counts_dev, self.finish = counts_dev.reshape(-1), finish


def _resolve(*pending):
    flat = torch.cat([p.counts_dev.to(torch.int64) for p in pending]).tolist()              # the one host sync
    _run_deferred_checks()
    out, o = [], 0
    for p in pending:
        n = p.counts_dev.numel()
        out.append(p.finish(flat[o:o + n]))
        o += n
    return out


def _filter_nms_batched_begin(boxes, scores, idxs, valid, nms_thresh, top_n):
    """The tail of torchvision's per-image loops -- drop the filtered boxes, ``batched_nms`` per category, keep the best
    ``top_n`` -- for all images at once and without a host sync per image.  boxes [B, M, 4], scores / idxs / valid [B, M].
    Instead of compacting each image (a ``nonzero`` + sync), filtered boxes get score -inf and sort to the end of the row;
    the stable descending sort orders the surviving boxes exactly as the compacted per-image sort would, the coordinate
    offset uses the maximum over the surviving boxes only, and hd_nms takes the survivor count from device memory.
    Resolves to per-image tuples (boxes, scores, idxs), identical to the torchvision loop."""
    B, M = scores.shape
    neg_inf = float("-inf")
    max_coord = boxes.masked_fill(~valid[..., None], neg_inf).amax(dim=(1, 2))                     # boxes.max() of the survivors
    offsets = idxs.to(boxes) * (max_coord + 1)[:, None]
    boxes_for_nms = boxes + offsets[..., None]
    order = torch.sort(scores.masked_fill(~valid, neg_inf), dim=1, descending=True, stable=True)[1]
    gidx = order[..., None].expand(-1, -1, 4)
    sorted_for_nms = torch.gather(boxes_for_nms, 1, gidx).contiguous()
    counts = valid.sum(1, dtype=torch.int32)
    keep = ops.nms_sorted_flat(sorted_for_nms.view(-1, 4), [i * M for i in range(B + 1)], nms_thresh, counts=counts).view(B, M)
    sel = keep & (keep.cumsum(1) <= top_n)

    def finish(n_sel):
        # sel is row-major, so the flat positions of its set bits are the per-image results back to back; their number is
        # known on the host now, so no boolean-mask indexing (each would be a nonzero + host sync)
        pos = torch.nonzero_static(sel.reshape(-1), size=sum(n_sel))[:, 0]
        src = (order + torch.arange(B, device=order.device)[:, None] * M).reshape(-1)[pos]      # flat index into [B*M]
        out_boxes = boxes.reshape(-1, 4)[src].split(n_sel)
        out_scores = scores.reshape(-1)[src].split(n_sel)
        out_idxs = idxs.reshape(-1)[src].split(n_sel)
        return out_boxes, out_scores, out_idxs
    return _Pending(sel.sum(1), finish)


def _filter_nms_batched(boxes, scores, idxs, valid, nms_thresh, top_n):
    return _resolve(_filter_nms_batched_begin(boxes, scores, idxs, valid, nms_thresh, top_n))[0]


def filter_proposals_batched_begin(rpn, proposals, objectness, image_shapes, num_anchors_per_level):
    """torchvision ``RegionProposalNetwork.filter_proposals`` (TV models/detection/rpn.py:242-295) with the per-image loop
    (clip, small-box / score filters, per-level NMS, top-n) done for the whole batch: ~30 launches and one host sync
    instead of ~25 launches and 3 syncs per image.  Results are identical (tests/test_modules_gpu.py)."""
    num_images = proposals.shape[0]
    device = proposals.device
    objectness = objectness.detach().reshape(num_images, -1)
    levels = torch.cat([torch.full((n,), idx, dtype=torch.int64, device=device) for idx, n in enumerate(num_anchors_per_level)], 0)
    levels = levels.reshape(1, -1).expand_as(objectness)
    top_n_idx = rpn._get_top_n_idx(objectness, num_anchors_per_level)
    batch_idx = torch.arange(num_images, device=device)[:, None]
    objectness = objectness[batch_idx, top_n_idx]
    levels = levels[batch_idx, top_n_idx]
    proposals = proposals[batch_idx, top_n_idx]
    scores = torch.sigmoid(objectness)
    with torch.no_grad():
        boxes = _clip_boxes_batched(proposals, image_shapes)
        ws, hs = boxes[..., 2] - boxes[..., 0], boxes[..., 3] - boxes[..., 1]
        valid = (ws >= rpn.min_size) & (hs >= rpn.min_size) & (scores >= rpn.score_thresh)
        pend = _filter_nms_batched_begin(boxes, scores, levels, valid, rpn.nms_thresh, rpn.post_nms_top_n())
    inner = pend.finish
    pend.finish = lambda n_sel: (lambda r: (list(r[0]), list(r[1])))(inner(n_sel))
    return pend


class _StaticProposals:
    """Proposals of the batch in a fixed-shape layout: ``boxes`` [B, T, 4] (T = post-NMS top-n; image b's proposals are
    rows 0 .. counts[b]-1 in torchvision's order, the rest zeros) and ``counts`` [B] on the DEVICE -- the host never learns
    how many proposals survived, so nothing waits for the proposal filter."""

    def __init__(self, boxes, counts):
        self.boxes, self.counts = boxes, counts


def _filter_nms_static(boxes, scores, idxs, valid, nms_thresh, top_n, group_sizes=None):
    """``_filter_nms_batched_begin`` with a fixed-shape result: the kept boxes of every image, in the same order, scattered to
    the front of a [B, top_n, 4] tensor; returns it with the per-image counts (device).

    ``group_sizes`` (host ints summing to M): the columns are laid out category by category (``idxs`` constant within a group)
    and every group is already in descending score order (the RPN's per-level top-k).  Boxes of different categories never
    suppress each other, so each (image, category) pair is then its own NMS problem -- 40 short problems running side by side
    instead of 8 long ones (the scan over a problem is sequential) -- on the SAME offset coordinates torchvision's
    ``batched_nms`` feeds its kernel, with the filtered boxes masked instead of sorted to the end.  Same keep set."""
    B, M = scores.shape
    neg_inf = float("-inf")
    max_coord = boxes.masked_fill(~valid[..., None], neg_inf).amax(dim=(1, 2))
    offsets = idxs.to(boxes) * (max_coord + 1)[:, None]
    boxes_for_nms = boxes + offsets[..., None]
    order = torch.sort(scores.masked_fill(~valid, neg_inf), dim=1, descending=True, stable=True)[1]
    gidx = order[..., None].expand(-1, -1, 4)
    if group_sizes is not None and sum(group_sizes) == M and len(group_sizes) * B <= 64:
        bounds = [0]
        for b in range(B):
            for g in group_sizes:
                bounds.append(bounds[-1] + g)
        keep0 = ops.nms_sorted_flat(boxes_for_nms.reshape(-1, 4).contiguous(), bounds, nms_thresh, valid=valid.reshape(-1).contiguous())
        keep = torch.gather(keep0.view(B, M), 1, order)
    else:
        sorted_for_nms = torch.gather(boxes_for_nms, 1, gidx).contiguous()
        counts = valid.sum(1, dtype=torch.int32)
        keep = ops.nms_sorted_flat(sorted_for_nms.view(-1, 4), [i * M for i in range(B + 1)], nms_thresh, counts=counts).view(B, M)
    csum = keep.cumsum(1)
    sel = keep & (csum <= top_n)
    dest = torch.where(sel, csum - 1, csum.new_full((), top_n))              # unselected boxes go to a dump column
    out = boxes.new_zeros(B, top_n + 1, 4)
    out.scatter_(1, dest[..., None].expand(-1, -1, 4), torch.gather(boxes, 1, gidx))
    return out[:, :top_n].contiguous(), sel.sum(1)


def filter_proposals_static(rpn, proposals, objectness, image_shapes, num_anchors_per_level):
    """``filter_proposals_batched`` without the host read of the survivor counts: returns a ``_StaticProposals``."""
    num_images = proposals.shape[0]
    device = proposals.device
    objectness = objectness.detach().reshape(num_images, -1)
    levels = torch.cat([torch.full((n,), idx, dtype=torch.int64, device=device) for idx, n in enumerate(num_anchors_per_level)], 0)
    levels = levels.reshape(1, -1).expand_as(objectness)
    top_n_idx = rpn._get_top_n_idx(objectness, num_anchors_per_level)
    batch_idx = torch.arange(num_images, device=device)[:, None]
    objectness = objectness[batch_idx, top_n_idx]
    levels = levels[batch_idx, top_n_idx]
    proposals = proposals[batch_idx, top_n_idx]
    scores = torch.sigmoid(objectness)
    with torch.no_grad():
        boxes = _clip_boxes_batched(proposals, image_shapes)
        ws, hs = boxes[..., 2] - boxes[..., 0], boxes[..., 3] - boxes[..., 1]
        valid = (ws >= rpn.min_size) & (hs >= rpn.min_size) & (scores >= rpn.score_thresh)
        per_level = [min(rpn.pre_nms_top_n(), n) for n in num_anchors_per_level]       # _get_top_n_idx: top-k per level, level-major
        out, n = _filter_nms_static(boxes, scores, levels, valid, rpn.nms_thresh, rpn.post_nms_top_n(),
                                    group_sizes=per_level if PER_LEVEL_NMS else None)
    return _StaticProposals(out, n)


_LEVEL_IDS = {}


def filter_proposals_static_fused(rpn, deltas, anchors, objectness, image_shapes, num_anchors_per_level):
    """``filter_proposals_static`` from the raw box deltas: the per-level top-k is taken on the objectness logits first, and only
    those candidates are decoded / clipped / tested, by one launch (ops.rpn_decode_selected) -- torchvision decodes all anchors
    and then gathers; the operations are element-wise, so the selected rows are bit-identical.  Needs one anchor set shared by
    the images of the batch and one image size (what the detector transform produces)."""
    num_images = len(anchors)
    device = deltas.device
    objectness = objectness.detach().reshape(num_images, -1)
    top_n_idx = rpn._get_top_n_idx(objectness, num_anchors_per_level).contiguous()
    per_level = [min(rpn.pre_nms_top_n(), n) for n in num_anchors_per_level]
    key = (tuple(per_level), num_images, str(device))
    levels = _LEVEL_IDS.get(key)
    if levels is None:
        if len(_LEVEL_IDS) > 8:
            _LEVEL_IDS.clear()
        one = torch.cat([torch.full((n,), i, dtype=torch.int64, device=device) for i, n in enumerate(per_level)], 0)
        levels = _LEVEL_IDS[key] = one.reshape(1, -1).expand(num_images, -1).contiguous()
    coder = rpn.box_coder
    with torch.no_grad():
        boxes, scores, valid = ops.rpn_decode_selected(objectness.contiguous(), deltas.detach().contiguous(), anchors[0].contiguous(),
                                                       top_n_idx, coder.weights, coder.bbox_xform_clip, image_shapes[0], rpn.min_size,
                                                       rpn.score_thresh)
        out, n = _filter_nms_static(boxes, scores, levels, valid, rpn.nms_thresh, rpn.post_nms_top_n(),
                                    group_sizes=per_level if PER_LEVEL_NMS else None)
    return _StaticProposals(out, n)


def filter_proposals_batched(rpn, proposals, objectness, image_shapes, num_anchors_per_level):
    return _resolve(filter_proposals_batched_begin(rpn, proposals, objectness, image_shapes, num_anchors_per_level))[0]


def postprocess_detections_batched(roi_heads, class_logits, box_regression, proposals, image_shapes):
    return _resolve(postprocess_detections_batched_begin(roi_heads, class_logits, box_regression, proposals, image_shapes))[0]


def postprocess_detections_batched_begin(roi_heads, class_logits, box_regression, proposals, image_shapes):
    """torchvision ``RoIHeads.postprocess_detections`` (TV models/detection/roi_heads.py:668-727) for the whole batch at
    once (see filter_proposals_batched); identical boxes / scores / labels."""
    device = class_logits.device
    num_classes = class_logits.shape[-1]
    boxes_per_image = [b.shape[0] for b in proposals]
    pred_boxes = _decode(roi_heads.box_coder, box_regression, proposals)
    pred_scores = F.softmax(class_logits, -1)
    B, n_max = len(boxes_per_image), max(boxes_per_image)
    if all(n == n_max for n in boxes_per_image):
        boxes = pred_boxes.view(B, n_max, num_classes, 4)
        scores = pred_scores.view(B, n_max, num_classes)
        present = None
    else:                                              # ragged: pad every image to the longest one
        boxes = pred_boxes.new_zeros(B, n_max, num_classes, 4)
        scores = pred_scores.new_zeros(B, n_max, num_classes)
        present = torch.zeros(B, n_max, dtype=torch.bool, device=device)
        for i, (pb, ps) in enumerate(zip(pred_boxes.split(boxes_per_image, 0), pred_scores.split(boxes_per_image, 0))):
            boxes[i, :pb.shape[0]], scores[i, :ps.shape[0]], present[i, :pb.shape[0]] = pb, ps, True
    boxes = _clip_boxes_batched(boxes, image_shapes)
    labels = torch.arange(num_classes, device=device).view(1, 1, -1).expand_as(scores)
    boxes, scores, labels = boxes[:, :, 1:], scores[:, :, 1:], labels[:, :, 1:]
    boxes, scores, labels = boxes.reshape(B, -1, 4), scores.reshape(B, -1), labels.reshape(B, -1)
    ws, hs = boxes[..., 2] - boxes[..., 0], boxes[..., 3] - boxes[..., 1]
    valid = (scores > roi_heads.score_thresh) & (ws >= 1e-2) & (hs >= 1e-2)
    if present is not None:
        valid = valid & present[:, :, None].expand(-1, -1, num_classes - 1).reshape(B, -1)
    pend = _filter_nms_batched_begin(boxes, scores, labels, valid, roi_heads.nms_thresh, roi_heads.detections_per_img)
    inner = pend.finish
    pend.finish = lambda n_sel: (lambda r: (list(r[0]), list(r[1]), list(r[2])))(inner(n_sel))
    return pend


# ---- whole-batch target assignment and sampling (torchvision loops one image at a time: ~15 launches + 2-3 syncs each) ----

def _pad_rows(rows, width, fill=0, counts=None):
    """List of [n_i, ...] tensors -> ([B, width, ...] padded with ``fill``, [B, width] bool presence mask).  With ``counts``
    (rows per image), ``rows`` may hold several pieces per image, image-major.  The row counts are host-known shapes: the
    destination row of every source row is computed on the host, so the padding is one concatenation + one index_copy for
    the whole batch (instead of a copy per image)."""
    if counts is None:
        counts = [int(r.shape[0]) for r in rows]
    B = len(counts)
    ref = rows[0]
    if len(rows) == B and all(c == width for c in counts):
        return torch.stack(rows), torch.ones(B, width, dtype=torch.bool, device=ref.device)
    out = ref.new_full((B * width,) + tuple(ref.shape[1:]), fill)
    present = torch.zeros(B * width, dtype=torch.bool, device=ref.device)
    if sum(counts) > 0:
        dst = np.concatenate([np.arange(b * width, b * width + c, dtype=np.int64) for b, c in enumerate(counts)])
        dst = torch.from_numpy(dst).to(ref.device, non_blocking=True)
        out.index_copy_(0, dst, torch.cat(rows))
        present.index_fill_(0, dst, True)
    return out.view((B, width) + tuple(ref.shape[1:])), present.view(B, width)


_GT_PAD = {"key": None, "val": None}


def _padded_gt(targets, dtype):
    """Ground-truth boxes / labels of the batch padded to [B, G, 4] / [B, G] (+ presence mask); the RPN and the RoI heads
    both need them, so the last result is kept (keyed by the identity and version of the target tensors)."""
    key = (dtype,) + tuple((id(t["boxes"]), t["boxes"]._version, id(t["labels"]), t["labels"]._version) for t in targets)
    if _GT_PAD["key"] != key:
        boxes = [t["boxes"].to(dtype) for t in targets]
        G = max(1, max(int(g.shape[0]) for g in boxes))
        gt, present = _pad_rows(boxes, G)
        gl, _ = _pad_rows([t["labels"] for t in targets], G)
        _GT_PAD["key"], _GT_PAD["val"] = key, (gt, present, gl, [(t["boxes"], t["labels"]) for t in targets])   # (keeps the ids alive)
    return _GT_PAD["val"][:3]


def _match_batched(matcher, gt_boxes, gt_present, boxes):
    """``det_utils.Matcher`` (TV models/detection/_utils.py) on ``box_iou(gt, boxes)`` for the whole batch.  gt_boxes
    [B, G, 4] padded (gt_present marks real rows), boxes [B, N, 4] -> matches [B, N] int64 (gt index, -1 below the low
    threshold, -2 between the thresholds).  Padded gt rows get quality -1, so they never win the arg-max, and they are
    excluded from the low-quality rule; an image without ground truth comes out all -1 (background), which is what
    torchvision's special case produces."""
    from torchvision.ops import boxes as box_ops
    mq = box_ops.box_iou(gt_boxes, boxes)                                    # [B, G, N]
    mq = mq.masked_fill(~gt_present[..., None], -1.0)
    matched_vals, matches = mq.max(dim=1)
    all_matches = matches.clone() if matcher.allow_low_quality_matches else None
    below = matched_vals < matcher.low_threshold
    between = (matched_vals >= matcher.low_threshold) & (matched_vals < matcher.high_threshold)
    matches = torch.where(below, matches.new_full((), matcher.BELOW_LOW_THRESHOLD), matches)
    matches = torch.where(between, matches.new_full((), matcher.BETWEEN_THRESHOLDS), matches)
    if matcher.allow_low_quality_matches:
        highest = mq.max(dim=2)[0]                                          # best quality of every gt
        update = ((mq == highest[..., None]) & gt_present[..., None]).any(dim=1)
        matches = torch.where(update, all_matches, matches)
    return matches


class _Samples:
    """Result of the batched sampler: flat indices into the [B * N] label array (image-major, in the order drawn within an
    image) of the sampled positives / negatives, and their per-image counts (host ints)."""

    def __init__(self, pos, neg, n_pos, n_neg):
        self.pos, self.neg, self.n_pos, self.n_neg = pos, neg, n_pos, n_neg

    def tensors(self):
        return [self.pos, self.neg]


def _gather_perms(nz, perms, starts, width):
    """nz: nonzero_static rows (image, column) of a [B, N] mask; perms[b] indexes the entries of image b (which start at row
    starts[b]).  Returns the selected entries as flat indices image * width + column -- one gather for the whole batch."""
    if len(perms) == 1:
        rows = perms[0] + starts[0] if starts[0] else perms[0]
    else:
        rows = torch.cat(torch._foreach_add(perms, starts))
    sel = nz[rows]
    return sel[:, 0] * width + sel[:, 1]


def _ascending(idx, total):
    """``torch.sort(idx)[0]`` for UNIQUE flat indices < total (sampled anchor / proposal positions) as mask + count-known
    ``nonzero``: 4 launches instead of the ~10 of a 64-bit radix sort (45 onesweep launches per step came from three such sorts)."""
    mask = torch.zeros(total, dtype=torch.bool, device=idx.device)
    mask[idx] = True
    return torch.nonzero_static(mask, size=idx.numel())[:, 0]


def _sample_batched_begin(sampler, labels):
    """``det_utils.BalancedPositiveNegativeSampler`` for labels [B, N] (>= 1 positive, 0 negative, -1 ignored / padding).
    Resolves to a ``_Samples``.  The two ``torch.randperm`` calls per image are issued with the same sizes and in the same
    order as torchvision's loop, so the CUDA generator is consumed identically and the samples are the same; everything
    else (counts, index lists, the gathers) is computed once for the batch."""
    B, N = labels.shape
    pos_mask, neg_mask = labels >= 1, labels == 0

    def finish(cnt):
        n_pos_all, n_neg_all = cnt[:B], cnt[B:]
        pos_nz = torch.nonzero_static(pos_mask, size=sum(n_pos_all))
        neg_nz = torch.nonzero_static(neg_mask, size=sum(n_neg_all))
        perms_p, perms_n, starts_p, starts_n, num_p, num_n, po, no = [], [], [], [], [], [], 0, 0
        for b in range(B):
            num_pos = min(n_pos_all[b], int(sampler.batch_size_per_image * sampler.positive_fraction))
            num_neg = min(n_neg_all[b], sampler.batch_size_per_image - num_pos)
            perms_p.append(torch.randperm(n_pos_all[b], device=labels.device)[:num_pos])
            perms_n.append(torch.randperm(n_neg_all[b], device=labels.device)[:num_neg])
            starts_p.append(po)
            starts_n.append(no)
            num_p.append(num_pos)
            num_n.append(num_neg)
            po, no = po + n_pos_all[b], no + n_neg_all[b]
        return _Samples(_gather_perms(pos_nz, perms_p, starts_p, N), _gather_perms(neg_nz, perms_n, starts_n, N), num_p, num_n)
    return _Pending(torch.cat([pos_mask.sum(1), neg_mask.sum(1)]), finish)


def _sample_batched(sampler, labels):
    return _resolve(_sample_batched_begin(sampler, labels))[0]


def assign_targets_to_anchors_batched(rpn, anchors, targets):
    """``RegionProposalNetwork.assign_targets_to_anchors`` (TV rpn.py) for the whole batch; returns [B, A] float labels
    (1 / 0 / -1) and [B, A, 4] matched boxes (identical values; the per-image lists are rows of these)."""
    A = torch.stack(anchors)                                                 # every image has the same anchor count
    gt, present, _ = _padded_gt(targets, targets[0]["boxes"].dtype)
    matches = _match_batched(rpn.proposal_matcher, gt, present, A)
    matched_gt = torch.gather(gt, 1, matches.clamp(min=0)[..., None].expand(-1, -1, 4))
    labels = (matches >= 0).to(torch.float32)
    labels = torch.where(matches == rpn.proposal_matcher.BETWEEN_THRESHOLDS, labels.new_full((), -1.0), labels)
    return labels, matched_gt.to(torch.float32) if matched_gt.dtype != torch.float32 else matched_gt


def rpn_compute_loss_batched(rpn, objectness, pred_bbox_deltas, labels, regression_targets, samples=None):
    """``RegionProposalNetwork.compute_loss`` (TV rpn.py) with labels [B, A] / regression_targets [B*A, 4] from the batched
    assignment.  Same samples (see _sample_batched), same reductions."""
    B, A = labels.shape
    if samples is None:
        samples = _sample_batched(rpn.fg_bg_sampler, labels)
    pos = _ascending(samples.pos, B * A)                                                 # == where(cat(pos masks))
    neg = _ascending(samples.neg, B * A)
    sampled = torch.cat([pos, neg], dim=0)
    objectness = objectness.flatten()
    labels = labels.reshape(-1)
    box_loss = F.smooth_l1_loss(pred_bbox_deltas[pos], regression_targets[pos], beta=1 / 9, reduction="sum") / (sampled.numel())
    objectness_loss = F.binary_cross_entropy_with_logits(objectness[sampled], labels[sampled])
    return objectness_loss, box_loss


def select_training_samples_batched(roi_heads, proposals, targets, return_num_pos=False):
    """``RoIHeads.select_training_samples`` (TV roi_heads.py: add_gt_proposals, assign_targets_to_proposals, subsample, the
    per-image gathers and BoxCoder.encode) for the whole batch.  Returns the same four per-image lists."""
    roi_heads.check_targets(targets)
    dtype, device = proposals[0].dtype, proposals[0].device
    B = len(proposals)
    gt_boxes = [t["boxes"].to(dtype) for t in targets]
    gt, present, gl = _padded_gt(targets, dtype)
    G = gt.shape[1]
    per_image_rows = [int(p.shape[0]) + int(g.shape[0]) for p, g in zip(proposals, gt_boxes)]
    N = max(per_image_rows)
    P, p_present = _pad_rows([x for pg in zip(proposals, gt_boxes) for x in pg], N, counts=per_image_rows)   # add_gt_proposals
    matches = _match_batched(roi_heads.proposal_matcher, gt, present, P)
    clamped = matches.clamp(min=0)
    labels = torch.gather(gl, 1, clamped).to(torch.int64)
    labels = torch.where(matches == roi_heads.proposal_matcher.BELOW_LOW_THRESHOLD, labels.new_zeros(()), labels)
    labels = torch.where(matches == roi_heads.proposal_matcher.BETWEEN_THRESHOLDS, labels.new_full((), -1), labels)
    labels = torch.where(p_present, labels, labels.new_full((), -1))        # padding is ignored by the sampler
    samples = _sample_batched(roi_heads.fg_bg_sampler, labels)
    per_image = [p + n for p, n in zip(samples.n_pos, samples.n_neg)]
    flat = _ascending(torch.cat((samples.pos, samples.neg)), B * N)             # == where(pos | neg) per image, back to back
    out_props = P.view(-1, 4)[flat]
    out_labels = labels.view(-1)[flat]
    out_matched = clamped.view(-1)[flat]
    img_of = torch.div(flat, N, rounding_mode="floor")
    matched_gt = gt.view(-1, 4)[img_of * G + out_matched]
    regression_targets = _encode_single(roi_heads.box_coder, matched_gt, out_props)
    out = (list(out_props.split(per_image)), list(out_matched.split(per_image)), list(out_labels.split(per_image)),
           list(regression_targets.split(per_image)))
    if return_num_pos:                                   # sampled foreground boxes == entries with label > 0 (host-known)
        return out + (sum(samples.n_pos),)
    return out


def _sampled_rows(sampled, counts, per_image):
    """Flat positions (ascending = image-major, ascending column: torchvision's ``where(pos | neg)`` lists back to back) of the
    drawn entries of ``sampled`` [B, N], as a FIXED-size list of B * per_image rows; rows past the drawn total (an image with
    fewer candidates than ``per_image``) point at entry 0 and are flagged invalid.  No host read."""
    B, N = sampled.shape
    S = B * per_image
    flat = torch.nonzero_static(sampled.view(-1), size=S, fill_value=0)[:, 0]
    n_drawn = counts[:, 2:4].sum()
    valid = torch.arange(S, device=sampled.device) < n_drawn
    return flat, valid, n_drawn


def rpn_targets_static(rpn, anchors, targets):
    """Anchor labels [B, A] and regression targets [B * A, 4] of the batch: two launches (ops.rpn_assign_targets, bit-identical)
    when the images share one anchor set, else ``assign_targets_to_anchors_batched`` + ``encode_boxes``."""
    gt, present, _ = _padded_gt(targets, targets[0]["boxes"].dtype)
    m = rpn.proposal_matcher
    if (FUSED_RPN_TARGETS and gt.dtype == torch.float32 and anchors[0].dtype == torch.float32 and gt.shape[1] <= 64
            and all(a.data_ptr() == anchors[0].data_ptr() or a.shape == anchors[0].shape for a in anchors)):
        return ops.rpn_assign_targets(anchors[0].contiguous(), gt.contiguous(), present.contiguous(), m.low_threshold, m.high_threshold,
                                      m.allow_low_quality_matches, rpn.box_coder.weights)
    labels, matched_gt_boxes = assign_targets_to_anchors_batched(rpn, anchors, targets)
    return labels, _encode_single(rpn.box_coder, matched_gt_boxes.reshape(-1, 4), torch.cat(anchors, dim=0))


def rpn_compute_loss_static(rpn, objectness, pred_bbox_deltas, labels, regression_targets, sampled, counts):
    """``RegionProposalNetwork.compute_loss`` (TV rpn.py) on the device-side draw (ops.sample_balanced): the sampled anchors as
    a fixed-size row list + validity weights instead of index lists whose lengths the host would have to read.  Same samples,
    same per-element losses, same divisors; only the order of the fp32 sums differs."""
    B, A = labels.shape
    if (FUSED_DET_LOSSES and objectness.is_cuda and objectness.dtype == torch.float32 and labels.dtype == torch.float32
            and regression_targets.dtype == torch.float32 and pred_bbox_deltas.dtype == torch.float32):
        flat = torch.nonzero_static(sampled.view(-1), size=B * rpn.fg_bg_sampler.batch_size_per_image, fill_value=0)[:, 0]
        return _RPNLoss.apply(objectness.reshape(-1), pred_bbox_deltas.reshape(-1, 4), labels.reshape(-1).contiguous(),
                              regression_targets.reshape(-1, 4).contiguous(), flat, sampled, counts)
    flat, valid, n_drawn = _sampled_rows(sampled, counts, rpn.fg_bg_sampler.batch_size_per_image)
    denom = n_drawn.to(torch.float32)
    is_pos = (sampled.view(-1)[flat] == 1) & valid
    # (the regression target of a background anchor is meaningless -- -inf / NaN when its image has no ground truth -- and
    # torchvision never touches it: zero it before the element-wise loss instead of multiplying a NaN by 0)
    tgt = torch.where(is_pos[:, None], regression_targets[flat], regression_targets.new_zeros(()))
    box = F.smooth_l1_loss(pred_bbox_deltas[flat], tgt, beta=1 / 9, reduction="none").sum(1)
    box_loss = torch.where(is_pos, box, box.new_zeros(())).sum() / denom
    obj = F.binary_cross_entropy_with_logits(objectness.flatten()[flat], labels.reshape(-1)[flat].clamp(min=0), reduction="none")
    objectness_loss = torch.where(valid, obj, obj.new_zeros(())).sum() / denom
    return objectness_loss, box_loss


class _StaticSamples:
    """RoI-head training samples in a fixed-shape layout (B * batch_size_per_image rows, image-major; rows past ``n_drawn``
    are padding: label -100, which cross_entropy ignores)."""

    def __init__(self, proposals, image_of, labels, regression_targets, matched_idxs, valid, n_drawn, per_image):
        self.proposals, self.image_of, self.labels, self.regression_targets = proposals, image_of, labels, regression_targets
        self.matched_idxs, self.valid, self.n_drawn, self.per_image = matched_idxs, valid, n_drawn, per_image
        self.rois = self.levels = None            # RoI format + pyramid level of every row (fused path)


def select_training_samples_static(roi_heads, proposals, targets, level_mapper=None):
    """``RoIHeads.select_training_samples`` for a ``_StaticProposals`` batch, without any host read: ground truth appended
    behind the (padded) proposals of each image, padding rows ignored by the sampler, the draw made on the device
    (ops.sample_balanced: the same selection as torchvision's loop on the same generator state).

    With FUSED_ROI_TARGETS the element-wise chains before and after the draw are the two launches of csrc/roi_targets.cu
    (bit-identical results); ``level_mapper`` (the pooler's LevelMapper) then also yields the RoIs and their pyramid levels."""
    roi_heads.check_targets(targets)
    P0, n_props = proposals.boxes, proposals.counts
    dtype, device = P0.dtype, P0.device
    B, T = P0.shape[:2]
    gt, present, gl = _padded_gt(targets, dtype)
    G = gt.shape[1]
    N = T + G
    matcher, sampler = roi_heads.proposal_matcher, roi_heads.fg_bg_sampler
    fused = (FUSED_ROI_TARGETS and dtype == torch.float32 and G <= 64 and not matcher.allow_low_quality_matches
             and gl.dtype == torch.int64 and P0.is_contiguous())
    if fused:
        labels, clamped = ops.roi_match_labels(P0, n_props.to(torch.int64), gt.contiguous(), present.contiguous(), gl.contiguous(),
                                               matcher.low_threshold, matcher.high_threshold)
        sampled, counts = ops.sample_balanced(labels, sampler.batch_size_per_image, sampler.positive_fraction)
        flat = torch.nonzero_static(sampled.view(-1), size=B * sampler.batch_size_per_image, fill_value=0)[:, 0]
        lm = level_mapper if level_mapper is not None else _NO_LEVELS
        o = ops.roi_gather_samples(flat, counts, P0, gt.contiguous(), labels, clamped, roi_heads.box_coder.weights, lm)
        valid = torch.arange(flat.numel(), device=device) < o["n_drawn"]
        smp = _StaticSamples(o["proposals"], o["image_of"], o["labels"], o["regression_targets"], o["matched"], valid, o["n_drawn"],
                             o["per_image"])
        if level_mapper is not None:
            smp.rois, smp.levels = o["rois"], o["levels"]
        return smp
    P = torch.cat([P0, gt], dim=1)                                                       # add_gt_proposals
    p_present = torch.cat([torch.arange(T, device=device)[None, :] < n_props[:, None], present], dim=1)
    matches = _match_batched(matcher, gt, present, P)
    clamped = matches.clamp(min=0)
    labels = torch.gather(gl, 1, clamped).to(torch.int64)
    labels = torch.where(matches == matcher.BELOW_LOW_THRESHOLD, labels.new_zeros(()), labels)
    labels = torch.where(matches == matcher.BETWEEN_THRESHOLDS, labels.new_full((), -1), labels)
    labels = torch.where(p_present, labels, labels.new_full((), -1))
    sampled, counts = ops.sample_balanced(labels.contiguous(), sampler.batch_size_per_image, sampler.positive_fraction)
    flat, valid, n_drawn = _sampled_rows(sampled, counts, sampler.batch_size_per_image)
    out_props = P.reshape(-1, 4)[flat]
    out_labels = torch.where(valid, labels.view(-1)[flat], labels.new_full((), -100))
    out_matched = clamped.view(-1)[flat]
    img_of = torch.div(flat, N, rounding_mode="floor")
    matched_gt = gt.reshape(-1, 4)[img_of * G + out_matched]
    regression_targets = _encode_single(roi_heads.box_coder, matched_gt, out_props)
    return _StaticSamples(out_props, img_of, out_labels, regression_targets, out_matched, valid, n_drawn, counts[:, 2] + counts[:, 3])


class _NoLevels:
    s0, lvl0, eps, k_min, k_max = 224, 4, 1e-6, 0, 0


_NO_LEVELS = _NoLevels()


class _FastRCNNLoss(torch.autograd.Function):
    """fastrcnn_loss with value and gradient from one launch (ops.fastrcnn_loss); backward only scales the stored gradients."""

    @staticmethod
    def forward(ctx, class_logits, box_regression, labels, regression_targets):
        losses, g_logits, g_box = ops.fastrcnn_loss(class_logits.detach().contiguous(), box_regression.detach().contiguous(), labels,
                                                    regression_targets)
        ctx.save_for_backward(g_logits, g_box)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        g_logits, g_box = ctx.saved_tensors
        return g_logits * g_cls, g_box * g_reg, None, None


class _RPNLoss(torch.autograd.Function):
    """RegionProposalNetwork.compute_loss on the device-side draw, value and gradient from one launch (ops.rpn_loss)."""

    @staticmethod
    def forward(ctx, objectness, pred_bbox_deltas, labels, regression_targets, flat, sampled, counts):
        losses, g_obj, g_deltas = ops.rpn_loss(objectness.detach().contiguous(), pred_bbox_deltas.detach().contiguous(), labels,
                                               regression_targets, flat, sampled, counts)
        ctx.save_for_backward(g_obj, g_deltas)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g0, g1):
        g_obj, g_deltas = ctx.saved_tensors
        return g_obj * g0, g_deltas * g1, None, None, None, None, None


def fastrcnn_loss_masked(class_logits, box_regression, samples):
    """``torchvision.models.detection.roi_heads.fastrcnn_loss`` on a ``_StaticSamples`` batch: the foreground rows enter the
    box loss through a 0 / 1 weight instead of ``torch.where(labels > 0)`` (a host read); padding rows carry the label
    cross_entropy ignores.  Same per-row losses and divisors as torchvision."""
    labels = samples.labels
    if (FUSED_DET_LOSSES and class_logits.is_cuda and class_logits.dtype == torch.float32 and box_regression.dtype == torch.float32
            and samples.regression_targets.dtype == torch.float32):
        return _FastRCNNLoss.apply(class_logits, box_regression, labels.contiguous(), samples.regression_targets.contiguous())
    classification_loss = F.cross_entropy(class_logits, labels)                           # ignore_index = -100: the padding rows
    S = class_logits.shape[0]
    pos = labels > 0
    per_class = box_regression.reshape(S, box_regression.size(-1) // 4, 4)
    picked = per_class[torch.arange(S, device=labels.device), labels.clamp(min=0)]
    tgt = torch.where(pos[:, None], samples.regression_targets, samples.regression_targets.new_zeros(()))   # (see rpn_compute_loss_static)
    box = F.smooth_l1_loss(picked, tgt, beta=1 / 9, reduction="none").sum(1)
    box_loss = torch.where(pos, box, box.new_zeros(())).sum() / samples.n_drawn.to(box.dtype)
    return classification_loss, box_loss


def postprocess_detections_static_begin(roi_heads, class_logits, box_regression, samples, image_shapes, per_image_max):
    """``postprocess_detections_batched_begin`` for a ``_StaticSamples`` batch: image b's rows start at the exclusive prefix
    sum of the per-image counts (device), so they are gathered into a [B, per_image_max] layout with a presence mask."""
    device = class_logits.device
    num_classes = class_logits.shape[-1]
    B = len(image_shapes)
    S = class_logits.shape[0]
    pred_boxes = _decode(roi_heads.box_coder, box_regression, [samples.proposals])
    pred_scores = F.softmax(class_logits, -1)
    n = samples.per_image.to(torch.int64)
    start = n.cumsum(0) - n
    j = torch.arange(per_image_max, device=device)[None, :]
    src = (start[:, None] + j).clamp(max=S - 1)
    present = j < n[:, None]
    boxes = pred_boxes.reshape(S, num_classes, 4)[src]
    scores = pred_scores[src]
    boxes = _clip_boxes_batched(boxes, image_shapes)
    labels = torch.arange(num_classes, device=device).view(1, 1, -1).expand_as(scores)
    boxes, scores, labels = boxes[:, :, 1:], scores[:, :, 1:], labels[:, :, 1:]
    boxes, scores, labels = boxes.reshape(B, -1, 4), scores.reshape(B, -1), labels.reshape(B, -1)
    ws, hs = boxes[..., 2] - boxes[..., 0], boxes[..., 3] - boxes[..., 1]
    valid = (scores > roi_heads.score_thresh) & (ws >= 1e-2) & (hs >= 1e-2)
    valid = valid & present[:, :, None].expand(-1, -1, num_classes - 1).reshape(B, -1)
    pend = _filter_nms_batched_begin(boxes, scores, labels, valid, roi_heads.nms_thresh, roi_heads.detections_per_img)
    inner = pend.finish
    pend.finish = lambda n_sel: (lambda r: (list(r[0]), list(r[1]), list(r[2])))(inner(n_sel))
    return pend


def fastrcnn_loss_static(class_logits, box_regression, labels, regression_targets, num_pos):
    """``torchvision.models.detection.roi_heads.fastrcnn_loss`` with the foreground count passed in (known on the host
    from the sampler), so ``torch.where(labels > 0)`` needs no device->host sync.  Same indices, same reductions."""
    labels = torch.cat(labels, dim=0)
    regression_targets = torch.cat(regression_targets, dim=0)
    classification_loss = F.cross_entropy(class_logits, labels)
    sampled_pos_inds_subset = torch.nonzero_static(labels > 0, size=num_pos)[:, 0]
    labels_pos = labels[sampled_pos_inds_subset]
    N, num_classes = class_logits.shape
    box_regression = box_regression.reshape(N, box_regression.size(-1) // 4, 4)
    box_loss = F.smooth_l1_loss(box_regression[sampled_pos_inds_subset, labels_pos], regression_targets[sampled_pos_inds_subset],
                                beta=1 / 9, reduction="sum")
    box_loss = box_loss / labels.numel()
    return classification_loss, box_loss


class _RoIAlign(torch.autograd.Function):
    """RoIAlign of one FPN level: forward on hd_roi_align_fwd_nhwc (bit-identical to torchvision's op, which is the
    fallback), backward on hd_roi_align_bwd_nhwc (ops.roi_align_bwd)."""

    @staticmethod
    def forward(ctx, feat, rois, spatial_scale, output_size, sampling_ratio, feat_nhwc):
        ctx.save_for_backward(rois)
        ctx.cfg = (tuple(feat.shape), float(spatial_scale), int(sampling_ratio))
        if feat_nhwc is not None:            # channels-last copy of feat: coalesced gathers, bit-identical pooled features
            return ops.roi_align_fwd(feat_nhwc, rois.contiguous(), output_size, spatial_scale, sampling_ratio)
        return torch.ops.torchvision.roi_align(feat, rois, float(spatial_scale), int(output_size[0]), int(output_size[1]),
                                               int(sampling_ratio), False)

    @staticmethod
    def backward(ctx, grad):
        (rois,) = ctx.saved_tensors
        shape, spatial_scale, sampling_ratio = ctx.cfg
        return ops.roi_align_bwd(grad.contiguous(), rois.contiguous(), shape, spatial_scale, sampling_ratio), None, None, None, None, None


def _roi_align(feat, rois, output_size, spatial_scale, sampling_ratio):
    from torchvision.ops import roi_align
    c = feat.shape[1]
    if (ROI_ALIGN_BWD and feat.is_cuda and feat.dtype == torch.float32 and rois.dtype == torch.float32 and feat.requires_grad
            and 1 <= sampling_ratio <= 2 and output_size[0] * output_size[1] <= 49 and c % 4 == 0 and c <= 256 and 256 % (c // 4) == 0):
        nhwc = ops.nchw_to_nhwc_f32(feat.detach().contiguous()) if ROI_ALIGN_FWD and rois.shape[0] > 0 else None
        return _RoIAlign.apply(feat, rois, spatial_scale, tuple(output_size), sampling_ratio, nhwc)
    return roi_align(feat, rois, output_size=output_size, spatial_scale=spatial_scale, sampling_ratio=sampling_ratio)


class _MultiLevelRoIAlign(torch.autograd.Function):
    """All FPN levels of MultiScaleRoIAlign in one launch each way (ops.roi_align_ml_fwd / _bwd): the level of every
    RoI is read on the device, so no per-level index lists and no host sync."""

    @staticmethod
    def forward(ctx, rois, levels, scales, output_size, sampling_ratio, bf16, *feats):
        nhwc, cl = [], []
        for f in feats:
            v = f.detach().permute(0, 2, 3, 1)
            cl.append(v.is_contiguous())           # channels-last feature map (backbone.CHANNELS_LAST_FEATURES): used in place
            if bf16 is None:
                nhwc.append(v if cl[-1] else ops.nchw_to_nhwc_f32(f.detach().contiguous()))
        if bf16 is not None:                       # the backbone's own bf16 NHWC pyramid (opt-in, see ROI_ALIGN_BF16)
            nhwc = list(bf16)
        ctx.save_for_backward(rois, levels)
        ctx.cfg = ([tuple(f.shape) for f in feats], tuple(scales), int(sampling_ratio), cl)
        return ops.roi_align_ml_fwd(nhwc, scales, rois, levels, output_size, sampling_ratio)

    @staticmethod
    def backward(ctx, grad):
        rois, levels = ctx.saved_tensors
        shapes, scales, sampling_ratio, cl = ctx.cfg
        grads = ops.roi_align_ml_bwd(grad.contiguous(), rois, levels, shapes, scales, sampling_ratio, channels_last=cl)
        return (None, None, None, None, None, None) + tuple(grads)


def _ml_roi_align_ok(feats, output_size, sampling_ratio):
    c = feats[0].shape[1]
    return (ROI_ALIGN_FUSED_LEVELS and ROI_ALIGN_FWD and ROI_ALIGN_BWD and len(feats) <= 8
            and all(f.is_cuda and f.dtype == torch.float32 and f.shape[1] == c for f in feats)
            and 1 <= sampling_ratio <= 2 and output_size[0] * output_size[1] <= 49 and c % 4 == 0 and c <= 256 and 256 % (c // 4) == 0)


def multiscale_roi_align_one_sync(pooler, features, boxes, image_shapes):
    """``torchvision.ops.MultiScaleRoIAlign.forward`` (TV ops/poolers.py) with the per-level ``torch.where(levels == k)``
    (one host sync per FPN level) replaced by a stable sort of the level ids and one ``bincount`` read: the index lists
    are the same (ascending within a level), so the pooled features are identical."""
    from torchvision.ops import boxes as box_ops, poolers, roi_align
    x_filtered = poolers._filter_input(features, pooler.featmap_names)
    if pooler.scales is None or pooler.map_levels is None:
        pooler.scales, pooler.map_levels = poolers._setup_scales(x_filtered, image_shapes, pooler.canonical_scale, pooler.canonical_level)
    num_levels = len(x_filtered)
    if num_levels == 1:
        return pooler(features, boxes, image_shapes)
    if sum(len(b) for b in boxes) > 0 and _ml_roi_align_ok(x_filtered, pooler.output_size, pooler.sampling_ratio):    # no host sync at all
        # _convert_to_roi_format and LevelMapper.__call__ (TV ops/poolers.py) on the concatenated boxes: the same
        # element-wise operations in the same order, without the per-image launches
        lm = pooler.map_levels
        concat = torch.cat(boxes, dim=0)
        ids = np.repeat(np.arange(len(boxes), dtype=np.float32), [len(b) for b in boxes])
        ids = torch.from_numpy(ids).to(concat.device, non_blocking=True).to(concat.dtype)
        rois = torch.cat([ids[:, None], concat], dim=1)
        s = torch.sqrt(box_ops.box_area(concat))
        target_lvls = torch.floor(lm.lvl0 + torch.log2(s / lm.s0) + torch.tensor(lm.eps, dtype=s.dtype))
        target_lvls = torch.clamp(target_lvls, min=lm.k_min, max=lm.k_max)
        levels = (target_lvls.to(torch.int64) - lm.k_min).to(torch.int64)
        return _MultiLevelRoIAlign.apply(rois, levels, tuple(float(sc) for sc in pooler.scales),
                                         tuple(pooler.output_size), int(pooler.sampling_ratio), None, *x_filtered)
    rois = poolers._convert_to_roi_format(boxes)
    levels = pooler.map_levels(boxes)
    order = torch.sort(levels, stable=True)[1]
    counts = torch.bincount(levels, minlength=num_levels).tolist()                 # the one host sync
    result = torch.zeros((len(rois), x_filtered[0].shape[1]) + tuple(pooler.output_size), dtype=x_filtered[0].dtype,
                         device=x_filtered[0].device)
    o = 0
    for level, (feat, scale) in enumerate(zip(x_filtered, pooler.scales)):
        idx_in_level = order[o:o + counts[level]]
        o += counts[level]
        pooled = _roi_align(feat, rois[idx_in_level], pooler.output_size, scale, pooler.sampling_ratio)
        result[idx_in_level] = pooled.to(result.dtype)
    return result


def _pooler_setup(pooler, features, image_shapes):
    from torchvision.ops import poolers
    x_filtered = poolers._filter_input(features, pooler.featmap_names)
    if pooler.scales is None or pooler.map_levels is None:
        pooler.scales, pooler.map_levels = poolers._setup_scales(x_filtered, image_shapes, pooler.canonical_scale, pooler.canonical_level)
    return x_filtered


def multiscale_roi_align_static(pooler, features, boxes, image_of, image_shapes, rois=None, levels=None, bf16=None):
    """``MultiScaleRoIAlign.forward`` for boxes [S, 4] whose image index is a device tensor (``_StaticSamples``): the RoI format
    and the level mapper are the element-wise operations of TV ops/poolers.py on the concatenated boxes; all levels are pooled
    by one launch each way (``_MultiLevelRoIAlign``).  Requires ``_ml_roi_align_ok``."""
    from torchvision.ops import boxes as box_ops, poolers
    x_filtered = poolers._filter_input(features, pooler.featmap_names)
    if pooler.scales is None or pooler.map_levels is None:
        pooler.scales, pooler.map_levels = poolers._setup_scales(x_filtered, image_shapes, pooler.canonical_scale, pooler.canonical_level)
    if bf16 is not None and (len(bf16) != len(x_filtered) or any(tuple(b.shape) != (f.shape[0], f.shape[2], f.shape[3], f.shape[1])
                                                                 for b, f in zip(bf16, x_filtered))):
        bf16 = None
    if rois is not None and levels is not None:
        return _MultiLevelRoIAlign.apply(rois, levels, tuple(float(sc) for sc in pooler.scales), tuple(pooler.output_size),
                                         int(pooler.sampling_ratio), bf16, *x_filtered)
    rois = torch.cat([image_of.to(boxes.dtype)[:, None], boxes], dim=1)
    if len(x_filtered) == 1:
        levels = torch.zeros(boxes.shape[0], dtype=torch.int64, device=boxes.device)
    else:
        lm = pooler.map_levels
        s = torch.sqrt(box_ops.box_area(boxes))
        target_lvls = torch.floor(lm.lvl0 + torch.log2(s / lm.s0) + torch.tensor(lm.eps, dtype=s.dtype))
        target_lvls = torch.clamp(target_lvls, min=lm.k_min, max=lm.k_max)
        levels = (target_lvls.to(torch.int64) - lm.k_min).to(torch.int64)
    return _MultiLevelRoIAlign.apply(rois, levels, tuple(float(sc) for sc in pooler.scales), tuple(pooler.output_size),
                                     int(pooler.sampling_ratio), bf16, *x_filtered)


def _static_tail_ok(model, features):
    """The sync-free training tail (STATIC_TAIL) needs the whole-batch restatements, the multi-level RoIAlign kernels and
    torchvision's stock sampler / pooler types."""
    from torchvision.ops import MultiScaleRoIAlign, poolers
    from torchvision.models.detection._utils import BalancedPositiveNegativeSampler
    feats = list(features.values())
    if not (STATIC_TAIL and BATCHED_TAIL and EARLY_RPN_TARGETS and feats[0].is_cuda and feats[0].dtype == torch.float32):
        return False
    rh = model.roi_heads
    if type(rh.box_roi_pool) is not MultiScaleRoIAlign or type(rh.fg_bg_sampler) is not BalancedPositiveNegativeSampler:
        return False
    if type(model.rpn.fg_bg_sampler) is not BalancedPositiveNegativeSampler or rh.has_mask() or rh.has_keypoint():
        return False
    if max(rh.fg_bg_sampler.batch_size_per_image, model.rpn.fg_bg_sampler.batch_size_per_image) > 1024:
        return False
    x_filtered = poolers._filter_input(features, rh.box_roi_pool.featmap_names)
    return _ml_roi_align_ok(x_filtered, rh.box_roi_pool.output_size, rh.box_roi_pool.sampling_ratio)


def _batched_ok(n_boxes):
    return n_boxes <= ops.NMS_MAX_BOXES and n_boxes * 4 <= 100_000


def filter_proposals_concurrent(rpn, proposals, objectness, image_shapes, num_anchors_per_level):
    """torchvision ``RegionProposalNetwork.filter_proposals`` (TV models/detection/rpn.py:242-295), same operators in the
    same order per image -- but the per-image bodies (clip, small-box / score filters, batched NMS, top-n) are enqueued
    on one CUDA stream per image, so the single-block NMS kernels of the images overlap instead of running back to
    back.  No randomness is involved; results are identical to the sequential loop."""
    from torchvision.ops import boxes as box_ops
    num_images = proposals.shape[0]
    device = proposals.device
    objectness = objectness.detach().reshape(num_images, -1)
    levels = torch.cat([torch.full((n,), idx, dtype=torch.int64, device=device) for idx, n in enumerate(num_anchors_per_level)], 0)
    levels = levels.reshape(1, -1).expand_as(objectness)
    top_n_idx = rpn._get_top_n_idx(objectness, num_anchors_per_level)
    batch_idx = torch.arange(num_images, device=device)[:, None]
    objectness = objectness[batch_idx, top_n_idx]
    levels = levels[batch_idx, top_n_idx]
    proposals = proposals[batch_idx, top_n_idx]
    objectness_prob = torch.sigmoid(objectness)
    if device.type != "cuda" or num_images == 1:
        streams = [None] * num_images
    else:
        streams = _side_streams(device, num_images)
    main = torch.cuda.current_stream(device) if device.type == "cuda" else None
    def body(boxes, scores, lvl, img_shape, st):
        # torchvision's nms op ends with a device->host read of the keep count, so overlapping the images needs one
        # host thread per image as well (ATen releases the GIL while it waits)
        ctx = torch.cuda.stream(st) if st is not None else torch.autograd.profiler.record_function("filter_proposals")
        with torch.no_grad(), ctx:
            boxes = box_ops.clip_boxes_to_image(boxes, img_shape)
            keep = box_ops.remove_small_boxes(boxes, rpn.min_size)
            boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
            keep = torch.where(scores >= rpn.score_thresh)[0]
            boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
            keep = batched_nms(boxes, scores, lvl, rpn.nms_thresh)
            keep = keep[: rpn.post_nms_top_n()]
            boxes, scores = boxes[keep], scores[keep]
        if st is not None:
            boxes.record_stream(main)
            scores.record_stream(main)
        return boxes, scores

    work = list(zip(proposals, objectness_prob, levels, image_shapes, streams))
    if streams[0] is None:
        results = [body(*w) for w in work]
    else:
        for st in streams:
            st.wait_stream(main)
        results = list(_thread_pool(num_images).map(lambda w: body(*w), work))
        for st in streams:
            main.wait_stream(st)
    return [r[0] for r in results], [r[1] for r in results]


_POOL = None


def _thread_pool(n):
    global _POOL
    if _POOL is None or _POOL._max_workers < n:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(n, 8))
    return _POOL


import os as _os

# Host threads + one CUDA stream per image for the per-image loops (A/B switches: HD_CONCURRENT_NMS / HD_CONCURRENT_POSTPROCESS)
CONCURRENT_NMS = _os.environ.get("HD_CONCURRENT_NMS", "1") != "0"                   # proposal filtering
CONCURRENT_POSTPROCESS = _os.environ.get("HD_CONCURRENT_POSTPROCESS", "0") != "0"   # final detections
BATCHED_TAIL = _os.environ.get("HD_BATCHED_TAIL", "1") != "0"   # whole-batch proposal filter / detections post-processing
ROI_ALIGN_BWD = _os.environ.get("HD_ROI_ALIGN_BWD", "1") != "0"   # RoIAlign backward on hd_roi_align_bwd_nhwc
POSTPROCESS_SIDE_STREAM = _os.environ.get("HD_POST_SIDE", "1") == "1"   # deferred train-time detections on a side stream
EARLY_RPN_TARGETS = _os.environ.get("HD_EARLY_RPN", "1") == "1"   # anchor targets + sampler on a side stream under the backbone forward
ROI_ALIGN_FUSED_LEVELS = _os.environ.get("HD_ROI_FUSED", "1") == "1"   # all FPN levels in one launch, no host sync
ROI_ALIGN_FWD = _os.environ.get("HD_ROI_ALIGN_FWD", "1") != "0"   # ... and forward on hd_roi_align_fwd_nhwc (bit-identical)
DEFER_DETECTIONS = False    # set by HalluciDetTrainer.training_step: roi_heads_eval returns a DeferredDetections


class DeferredDetections:
    """Final detections whose per-image counts are still on their way to the host (pinned, non-blocking copy).
    ``resolve()`` waits for that copy only, assembles the per-image ``{"boxes", "labels", "scores"}`` dicts and applies the
    queued post-processing (the transform's rescaling to the original image size).  Used inside the train step, where the
    detections are a by-product (train_hallucidet.py:181 logs them): the step's critical path then has one host sync less."""

    def __init__(self, pending, stream=None):
        self._pending = pending
        self._stream = stream             # side stream the detections are computed on (None: the current stream)
        self._counts = torch.empty(pending.counts_dev.numel(), dtype=torch.int64, pin_memory=True)
        self._counts.copy_(pending.counts_dev.to(torch.int64), non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()
        self._post = []
        self._value = None

    def then(self, fn):
        self._post.append(fn)
        return self

    def resolve(self):
        if self._value is None:
            self._event.synchronize()
            _run_deferred_checks()                    # input checks queued by a sync-free step are evaluated here
            ops.DeviceRng.get(self._pending.counts_dev.device).sync_host()          # device-side sampler draws -> torch's generator
            side = self._stream
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()), torch.no_grad():
                boxes, scores, labels = self._pending.finish(self._counts.tolist())
                value = [{"boxes": boxes[i], "labels": labels[i], "scores": scores[i]} for i in range(len(boxes))]
                for fn in self._post:
                    value = fn(value)
            if side is not None:                      # hand the results over to the caller's stream
                main = torch.cuda.current_stream(side.device)
                main.wait_stream(side)
                for d in value:
                    for t in d.values():
                        if torch.is_tensor(t) and t.is_cuda:
                            t.record_stream(main)
            self._value, self._pending = value, None
        return self._value


class DeferredCall:
    """A by-product of the step (here: RetinaNet's detections, which feed no loss) computed LATER: ``resolve()`` runs ``fn`` on a
    side stream that only waits for what had been enqueued when the object was created.  The trainer resolves it after the
    backward pass and the optimizer have been enqueued, so the per-image / per-level loops of the post-processing (with their
    data-dependent host syncs) execute next to the backward pass instead of in front of it.  Same interface as
    DeferredDetections."""

    def __init__(self, fn, inputs):
        self._fn, self._value = fn, None
        self._stream = _side_streams(inputs[0].device, 2)[1]
        self._event = torch.cuda.Event()
        self._event.record()
        for t in inputs:
            t.record_stream(self._stream)

    def then(self, fn):
        inner = self._fn
        self._fn = lambda: fn(inner())
        return self

    def resolve(self):
        if self._value is None:
            side = self._stream
            side.wait_event(self._event)
            with torch.cuda.stream(side), torch.no_grad():
                value = self._fn()
                _run_deferred_checks()
                ops.DeviceRng.get(side.device).sync_host()        # device-side sampler draws -> torch's generator
            main = torch.cuda.current_stream(side.device)
            main.wait_stream(side)
            for d in value:
                for t in d.values():
                    if torch.is_tensor(t) and t.is_cuda:
                        t.record_stream(main)
            self._value, self._fn = value, None
        return self._value


GRAPH_PROPOSAL_FILTER = _os.environ.get("HD_GRAPH_PROPOSALS", "1") == "1"
# Training tail without device->host reads: fixed-shape proposals, device-side sampler draws (ops.sample_balanced), masked losses
STATIC_TAIL = _os.environ.get("HD_STATIC_TAIL", "1") == "1"
FUSED_ROI_TARGETS = _os.environ.get("HD_FUSED_ROI_TARGETS", "1") == "1"   # csrc/roi_targets.cu instead of ~80 element-wise launches
FUSED_DET_LOSSES = _os.environ.get("HD_FUSED_DET_LOSSES", "1") == "1"     # csrc/det_losses.cu: loss value + gradient in one launch
FLAT_RPN_PREDS = _os.environ.get("HD_FLAT_RPN_PREDS", "1") == "1"         # RPN head outputs flattened by one launch each way (heads.py)
# RoIAlign forward on the backbone's bf16 pyramid (half the L2 traffic, -0.06 ms).  OFF: the fp32 maps the backbone hands out
# are its un-rounded fp32 accumulators, not widened copies of the bf16 pyramid, so this would round the box head's input once
# more (loss_classifier moves by 5e-5 relative) for a 0.4 % gain.
ROI_ALIGN_BF16 = _os.environ.get("HD_ROI_ALIGN_BF16", "0") == "1"
STATIC_FEATURES = _os.environ.get("HD_STATIC_FEATURES", "1") == "1"       # feature maps = the backbone engine's own buffers (no clones)
FUSED_PROPOSAL_DECODE = _os.environ.get("HD_FUSED_PROPOSAL_DECODE", "1") == "1"   # decode only the per-level top-k (one launch)
FUSED_RPN_TARGETS = _os.environ.get("HD_FUSED_RPN_TARGETS", "1") == "1"     # anchor target assignment + encode as two launches
PER_LEVEL_NMS = _os.environ.get("HD_PER_LEVEL_NMS", "1") == "1"     # proposal NMS as (image, level) problems (see _filter_nms_static)
_STATIC_PROGRAMS = {}


class _StaticProgram:
    """A gradient-free, static-shape piece of the tail on static input buffers: run eagerly once (warm-up), captured into a
    CUDA graph on the second call, replayed afterwards.  The captured function's return value (tensors in the graph's memory
    pool, refreshed by every replay, and closures over them) is handed out each time."""

    def __init__(self):
        self.state, self.graph, self.out = None, None, None

    def run(self, fn):
        if self.state is None:
            self.state = "warm"
            return fn()
        if self.state == "warm":
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.out = fn()
            self.graph, self.state = g, "graph"
        self.graph.replay()
        return self.out


def _static_program(key):
    prog = _STATIC_PROGRAMS.get(key)
    if prog is None:
        if len(_STATIC_PROGRAMS) > 8:
            _STATIC_PROGRAMS.clear()
        prog = _STATIC_PROGRAMS[key] = _StaticProgram()
    return prog


_ANCHOR_CACHE = {}


def _anchors(model, images, features):
    """``AnchorGenerator.forward`` is a pure function of the image / feature-map shapes: cached per shape set, so that
    from the second step on the anchors do not queue behind the backbone forward on the stream (see rpn_eval)."""
    gen = model.rpn.anchor_generator
    key = (id(gen), tuple(images.tensors.shape), tuple(map(tuple, images.image_sizes)), tuple(tuple(f.shape) for f in features),
           features[0].dtype, features[0].device)
    hit = _ANCHOR_CACHE.get(key)
    fresh = hit is None
    if fresh:
        if len(_ANCHOR_CACHE) > 16:
            _ANCHOR_CACHE.clear()
        hit = _ANCHOR_CACHE[key] = (gen, gen(images, features))        # (the generator is kept alive so its id stays unique)
    return hit[1], fresh


FUSED_RPN_PREDICTORS = _os.environ.get("HD_FUSED_RPN_PRED", "1") == "1"
WHOLE_BATCH_RETINANET_LOSS = _os.environ.get("HD_RETINA_LOSS_BATCHED", "1") == "1"   # RetinaNet losses without per-image loop / host sync
B200_HEADS = _os.environ.get("HD_B200_HEADS", "1") == "1"       # frozen RPN / RetinaNet heads on the B200 conv kernels (heads.py)


def _rpn_head(head, features):
    """``RPNHead.forward`` (TV rpn.py:61-68) with the two frozen 1x1 predictors (objectness, box deltas) evaluated as ONE
    convolution over the concatenated output channels: the shared hidden map -- the largest activation of the tail -- is
    read once instead of twice, and its gradient is written once instead of being the sum of two full-size tensors."""
    cls, reg = head.cls_logits, head.bbox_pred
    ok = (FUSED_RPN_PREDICTORS and features[0].is_cuda and isinstance(cls, torch.nn.Conv2d) and isinstance(reg, torch.nn.Conv2d)
          and cls.kernel_size == reg.kernel_size == (1, 1) and cls.stride == reg.stride == (1, 1)
          and cls.padding == reg.padding == (0, 0) and cls.groups == reg.groups == 1
          and cls.bias is not None and reg.bias is not None
          and not any(p.requires_grad for p in (cls.weight, cls.bias, reg.weight, reg.bias)))
    if not ok:
        return head(features)
    weight, bias = torch.cat([cls.weight, reg.weight], 0), torch.cat([cls.bias, reg.bias], 0)
    a = cls.out_channels
    logits, bbox_reg = [], []
    for feature in features:
        y = F.conv2d(head.conv(feature), weight, bias)
        logits.append(y[:, :a])
        bbox_reg.append(y[:, a:])
    return logits, bbox_reg


def rpn_eval(model, images, features, targets, targets_event=None, static=False):
    """``RegionProposalNetwork.forward`` in training mode (TV rpn.py:337-388) as the reference's eval_forward uses it
    (src/utils/eval_forward_fasterrcnn.py:62-99).  ``targets_event``: CUDA event recorded once ``targets`` are final on
    the current stream; lets the anchor-target work start before the backbone forward has finished."""
    features = list(features.values())
    objectness, static_preds, flat_obj = None, False, None
    if B200_HEADS and features[0].is_cuda and isinstance(model.backbone, FrozenBackbone):
        # the frozen RPN head on the tcgen05 conv kernels, reading the backbone's bf16 pyramid (forward + input gradient only)
        bf16 = model.backbone.bf16_features()
        if bf16 is not None and len(bf16) == len(features) and heads.rpn_head_tower(model.rpn.head) is not None:
            heads.USE_CUDA_GRAPH = bool(model.backbone.use_cuda_graph)
            if FLAT_RPN_PREDS and features[0].dtype == torch.float32:
                # the head's outputs directly in torchvision's flattened per-anchor order (one launch each way)
                flat_obj, flat_deltas, flat_napl, static_preds = heads.rpn_head_forward_flat(model.rpn.head, features, bf16)
                objectness = flat_obj
            else:
                objectness, pred_bbox_deltas, static_preds = heads.rpn_head_forward(model.rpn.head, features, bf16, return_static=True)
    if objectness is None:
        objectness, pred_bbox_deltas = _rpn_head(model.rpn.head, features)
    batched = BATCHED_TAIL and features[0].is_cuda
    anchors, anchors_fresh = _anchors(model, images, features) if batched else (model.rpn.anchor_generator(images, features), True)
    num_images = len(anchors)
    if flat_obj is not None:
        num_anchors_per_level = flat_napl
    else:
        num_anchors_per_level = [o[0].shape[0] * o[0].shape[1] * o[0].shape[2] for o in objectness]
    pre_nms = sum(min(model.rpn.pre_nms_top_n(), n) for n in num_anchors_per_level)
    if targets is None:
        raise ValueError("targets should not be None")
    use_batched = (batched and (flat_obj if flat_obj is not None else objectness[0]).dtype == torch.float32 and _batched_ok(pre_nms)
                   and all(a.shape == anchors[0].shape for a in anchors))
    pend_boxes = None
    fused_decode = bool(static and use_batched and FUSED_PROPOSAL_DECODE and features[0].dtype == torch.float32
                        and anchors[0].dtype == torch.float32
                        and all(tuple(sz) == tuple(images.image_sizes[0]) for sz in images.image_sizes))
    if use_batched and static_preds and GRAPH_PROPOSAL_FILTER and not anchors_fresh:
        # The predictor outputs live in the static buffers of the head's CUDA-graph program and the anchors are cached: the whole
        # gradient-free, static-shape chain concat -> decode -> per-level top-k -> clip / filter -> sort -> NMS (~45 launches) is
        # captured once and replayed as ONE graph launch; only the final data-dependent gather stays eager (after the count read).
        if flat_obj is not None:
            obj_lv, del_lv = [flat_obj.detach()], [flat_deltas.detach()]
        else:
            obj_lv, del_lv = [o.detach() for o in objectness], [d.detach() for d in pred_bbox_deltas]
        image_sizes = images.image_sizes

        def program():
            o2, d2 = (obj_lv[0], del_lv[0]) if flat_obj is not None else concat_box_prediction_layers(obj_lv, del_lv)
            if static and fused_decode:
                return filter_proposals_static_fused(model.rpn, d2, anchors, o2, image_sizes, num_anchors_per_level)
            props = _decode(model.rpn.box_coder, d2, anchors).view(num_images, -1, 4)
            if static:
                return filter_proposals_static(model.rpn, props, o2, image_sizes, num_anchors_per_level)
            return filter_proposals_batched_begin(model.rpn, props, o2, image_sizes, num_anchors_per_level)

        key = (id(model.rpn), tuple(o.data_ptr() for o in obj_lv), anchors[0].data_ptr(), tuple(map(tuple, image_sizes)), bool(static))
        with torch.no_grad():
            pend_boxes = _static_program(key).run(program)
    if flat_obj is not None:
        objectness, pred_bbox_deltas = flat_obj, flat_deltas
    else:
        objectness, pred_bbox_deltas = concat_box_prediction_layers(objectness, pred_bbox_deltas)
    if pend_boxes is None and fused_decode:
        with torch.no_grad():
            pend_boxes = filter_proposals_static_fused(model.rpn, pred_bbox_deltas.detach(), anchors, objectness, images.image_sizes,
                                                       num_anchors_per_level)
    if pend_boxes is None:
        proposals = _decode(model.rpn.box_coder, pred_bbox_deltas.detach(), anchors)
        proposals = proposals.view(num_images, -1, 4)
    if use_batched:
        # The proposal filter (up to its NMS) is enqueued on the main stream.  Target assignment, box encoding and the
        # anchor sampler depend only on the anchors and the targets, not on the network: they run on a side stream as soon
        # as the targets exist -- i.e. underneath the backbone forward -- including the sampler's host read and its
        # randperm calls, which therefore leave the critical path (same calls in the same order: same CUDA generator use).
        static = static and EARLY_RPN_TARGETS
        if pend_boxes is None:
            if static:
                pend_boxes = filter_proposals_static(model.rpn, proposals, objectness, images.image_sizes, num_anchors_per_level)
            else:
                pend_boxes = filter_proposals_batched_begin(model.rpn, proposals, objectness, images.image_sizes, num_anchors_per_level)
        main = torch.cuda.current_stream(objectness.device)
        if static:
            # sync-free variant: the anchor sampler's draw stays on the device (same selection, see ops.sample_balanced), the loss
            # takes the sampled anchors as a fixed-size row list, and the proposals keep their fixed-shape layout
            side = _side_streams(objectness.device, 1)[0]
            if targets_event is None or anchors_fresh:
                targets_event = torch.cuda.Event()
                targets_event.record(main)
            side.wait_event(targets_event)
            sampler = model.rpn.fg_bg_sampler
            with torch.cuda.stream(side), torch.no_grad():
                labels, regression_targets = rpn_targets_static(model.rpn, anchors, targets)
                sampled, counts = ops.sample_balanced(labels.contiguous(), sampler.batch_size_per_image, sampler.positive_fraction)
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
            for t in [labels, regression_targets, sampled, counts] + list(_padded_gt(targets, targets[0]["boxes"].dtype)):
                t.record_stream(main)
            loss_objectness, loss_rpn_box_reg = rpn_compute_loss_static(model.rpn, objectness, pred_bbox_deltas, labels,
                                                                        regression_targets, sampled, counts)
            return pend_boxes, {"loss_objectness": loss_objectness, "loss_rpn_box_reg": loss_rpn_box_reg}
        if EARLY_RPN_TARGETS:
            side = _side_streams(objectness.device, 1)[0]
            if targets_event is None or anchors_fresh:
                # (anchors computed just now are queued on the main stream behind the backbone: wait for them as well)
                targets_event = torch.cuda.Event()
                targets_event.record(main)
            side.wait_event(targets_event)
            with torch.cuda.stream(side), torch.no_grad():
                labels, matched_gt_boxes = assign_targets_to_anchors_batched(model.rpn, anchors, targets)
                regression_targets = _encode_single(model.rpn.box_coder, matched_gt_boxes.reshape(-1, 4), torch.cat(anchors, dim=0))
                (samples,) = _resolve(_sample_batched_begin(model.rpn.fg_bg_sampler, labels))      # waits for the side stream only
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
            for t in [labels, regression_targets] + samples.tensors() + list(_padded_gt(targets, targets[0]["boxes"].dtype)):
                t.record_stream(main)
            # the loss does not need the proposals: enqueue it before waiting for their counts
            loss_objectness, loss_rpn_box_reg = rpn_compute_loss_batched(model.rpn, objectness, pred_bbox_deltas, labels,
                                                                         regression_targets, samples=samples)
            ((boxes, scores),) = _resolve(pend_boxes)
            return boxes, {"loss_objectness": loss_objectness, "loss_rpn_box_reg": loss_rpn_box_reg}
        else:
            with torch.no_grad():
                labels, matched_gt_boxes = assign_targets_to_anchors_batched(model.rpn, anchors, targets)
                regression_targets = _encode_single(model.rpn.box_coder, matched_gt_boxes.reshape(-1, 4), torch.cat(anchors, dim=0))
                pend_samples = _sample_batched_begin(model.rpn.fg_bg_sampler, labels)
            (boxes, scores), samples = _resolve(pend_boxes, pend_samples)
        loss_objectness, loss_rpn_box_reg = rpn_compute_loss_batched(model.rpn, objectness, pred_bbox_deltas, labels, regression_targets,
                                                                     samples=samples)
        return boxes, {"loss_objectness": loss_objectness, "loss_rpn_box_reg": loss_rpn_box_reg}
    if CONCURRENT_NMS and proposals.is_cuda:
        boxes, scores = filter_proposals_concurrent(model.rpn, proposals, objectness, images.image_sizes, num_anchors_per_level)
    else:
        boxes, scores = model.rpn.filter_proposals(proposals, objectness, images.image_sizes, num_anchors_per_level)
    labels, matched_gt_boxes = model.rpn.assign_targets_to_anchors(anchors, targets)
    regression_targets = model.rpn.box_coder.encode(matched_gt_boxes, anchors)
    loss_objectness, loss_rpn_box_reg = model.rpn.compute_loss(objectness, pred_bbox_deltas, labels, regression_targets)
    return boxes, {"loss_objectness": loss_objectness, "loss_rpn_box_reg": loss_rpn_box_reg}


def postprocess_detections_concurrent(roi_heads, class_logits, box_regression, proposals, image_shapes, threaded=True):
    """torchvision ``RoIHeads.postprocess_detections`` (TV models/detection/roi_heads.py:668-727), identical operators per
    image, with the per-image bodies (score filter, small-box filter, batched NMS, top-k) on one host thread + CUDA stream
    per image (see filter_proposals_concurrent).  No randomness; results equal the sequential loop."""
    from torchvision.ops import boxes as box_ops
    device = class_logits.device
    num_classes = class_logits.shape[-1]
    boxes_per_image = [b.shape[0] for b in proposals]
    pred_boxes = roi_heads.box_coder.decode(box_regression, proposals)
    pred_scores = F.softmax(class_logits, -1)
    pred_boxes_list = pred_boxes.split(boxes_per_image, 0)
    pred_scores_list = pred_scores.split(boxes_per_image, 0)
    n = len(boxes_per_image)
    main = torch.cuda.current_stream(device)
    streams = _side_streams(device, n)

    def body(boxes, scores, image_shape, st):
        with torch.no_grad(), torch.cuda.stream(st):
            boxes = box_ops.clip_boxes_to_image(boxes, image_shape)
            labels = torch.arange(num_classes, device=device).view(1, -1).expand_as(scores)
            boxes, scores, labels = boxes[:, 1:], scores[:, 1:], labels[:, 1:]
            boxes, scores, labels = boxes.reshape(-1, 4), scores.reshape(-1), labels.reshape(-1)
            inds = torch.where(scores > roi_heads.score_thresh)[0]
            boxes, scores, labels = boxes[inds], scores[inds], labels[inds]
            keep = box_ops.remove_small_boxes(boxes, min_size=1e-2)
            boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
            keep = batched_nms(boxes, scores, labels, roi_heads.nms_thresh)
            keep = keep[: roi_heads.detections_per_img]
            boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
        for t in (boxes, scores, labels):
            t.record_stream(main)
        return boxes, scores, labels

    work = list(zip(pred_boxes_list, pred_scores_list, image_shapes, streams))
    if not threaded:
        results = [body(b, s_, sh, main) for b, s_, sh, _ in work]
    else:
        for st in streams:
            st.wait_stream(main)
        results = list(_thread_pool(n).map(lambda w: body(*w), work))
        for st in streams:
            main.wait_stream(st)
    return [r[0] for r in results], [r[1] for r in results], [r[2] for r in results]


def _roi_heads_eval_static(model, features, proposals, image_shapes, targets):
    """``roi_heads_eval`` for ``_StaticProposals``: no device->host read between the backbone and the losses.  The detections
    (a by-product in the train step) are post-processed on a side stream and assembled when the caller resolves them."""
    rh = model.roi_heads
    x_filtered = _pooler_setup(rh.box_roi_pool, features, image_shapes)
    with torch.no_grad():
        samples = select_training_samples_static(rh, proposals, targets,
                                                 level_mapper=rh.box_roi_pool.map_levels if len(x_filtered) > 1 else None)
    bf16 = None
    if ROI_ALIGN_BF16 and isinstance(model.backbone, FrozenBackbone):
        pyramid = model.backbone.bf16_features()                       # same order as the feature dict (incl. the pooled level)
        if pyramid is not None and len(pyramid) == len(features):
            names = list(features.keys())
            bf16 = [pyramid[names.index(n)] for n in rh.box_roi_pool.featmap_names if n in names]
    box_features = multiscale_roi_align_static(rh.box_roi_pool, features, samples.proposals, samples.image_of, image_shapes,
                                               rois=samples.rois, levels=samples.levels, bf16=bf16)
    box_features = rh.box_head(box_features)
    class_logits, box_regression = rh.box_predictor(box_features)
    loss_classifier, loss_box_reg = fastrcnn_loss_masked(class_logits, box_regression, samples)
    losses = {"loss_classifier": loss_classifier, "loss_box_reg": loss_box_reg}
    per_image_max = rh.fg_bg_sampler.batch_size_per_image
    with torch.no_grad():
        cl, br = class_logits.detach(), box_regression.detach()
        if DEFER_DETECTIONS and POSTPROCESS_SIDE_STREAM:
            # the detections feed no loss: even ISSUING their post-processing (~100 small launches of host time) waits until
            # the caller has enqueued the backward pass; it then runs on a side stream next to it
            def late():
                boxes, scores, labels = _resolve(postprocess_detections_static_begin(rh, cl, br, samples, image_shapes, per_image_max))[0]
                return [{"boxes": boxes[i], "labels": labels[i], "scores": scores[i]} for i in range(len(boxes))]
            return DeferredCall(late, [cl, br, samples.proposals, samples.per_image]), losses
        pend = postprocess_detections_static_begin(rh, cl, br, samples, image_shapes, per_image_max)
        if DEFER_DETECTIONS:
            return DeferredDetections(pend), losses
        boxes, scores, labels = _resolve(pend)[0]
    ops.DeviceRng.get(cl.device).sync_host()
    return [{"boxes": boxes[i], "labels": labels[i], "scores": scores[i]} for i in range(len(boxes))], losses


def roi_heads_eval(model, features, proposals, image_shapes, targets=None, train_det=False):
    for t in targets:
        if t["boxes"].dtype not in (torch.float, torch.double, torch.half):
            raise TypeError(f"target boxes must of float type, instead got {t['boxes'].dtype}")
        if t["labels"].dtype != torch.int64:
            raise TypeError(f"target labels must of int64 type, instead got {t['labels'].dtype}")
    from torchvision.ops import MultiScaleRoIAlign
    if isinstance(proposals, _StaticProposals):
        return _roi_heads_eval_static(model, features, proposals, image_shapes, targets)
    num_pos = None
    if BATCHED_TAIL and proposals[0].is_cuda:
        with torch.no_grad():
            proposals, matched_idxs, labels, regression_targets, num_pos = select_training_samples_batched(
                model.roi_heads, proposals, targets, return_num_pos=True)
    else:
        proposals, matched_idxs, labels, regression_targets = model.roi_heads.select_training_samples(proposals, targets)
    if BATCHED_TAIL and proposals[0].is_cuda and type(model.roi_heads.box_roi_pool) is MultiScaleRoIAlign:
        box_features = multiscale_roi_align_one_sync(model.roi_heads.box_roi_pool, features, proposals, image_shapes)
    else:
        box_features = model.roi_heads.box_roi_pool(features, proposals, image_shapes)
    box_features = model.roi_heads.box_head(box_features)
    class_logits, box_regression = model.roi_heads.box_predictor(box_features)
    if num_pos is not None:
        loss_classifier, loss_box_reg = fastrcnn_loss_static(class_logits, box_regression, labels, regression_targets, num_pos)
    else:
        loss_classifier, loss_box_reg = fastrcnn_loss(class_logits, box_regression, labels, regression_targets)
    losses = {"loss_classifier": loss_classifier, "loss_box_reg": loss_box_reg}
    n_cand = max(p.shape[0] for p in proposals) * (class_logits.shape[-1] - 1)
    if BATCHED_TAIL and class_logits.is_cuda and class_logits.dtype == torch.float32 and _batched_ok(n_cand):
        with torch.no_grad():
            if DEFER_DETECTIONS and POSTPROCESS_SIDE_STREAM:
                # the detections feed no loss: they are computed on a side stream (underneath the backward pass), their
                # data-dependent sizes are read back asynchronously and the lists are assembled when the caller asks
                # (HalluciDetTrainer.training_step: after the backward pass and the optimizer step have been enqueued)
                main = torch.cuda.current_stream(class_logits.device)
                side = _side_streams(class_logits.device, 1)[0]
                side.wait_stream(main)
                cl, br = class_logits.detach(), box_regression.detach()
                for t in [cl, br] + list(proposals):
                    t.record_stream(side)
                with torch.cuda.stream(side):
                    pend = postprocess_detections_batched_begin(model.roi_heads, cl, br, proposals, image_shapes)
                    return DeferredDetections(pend, side), losses
            pend = postprocess_detections_batched_begin(model.roi_heads, class_logits.detach(), box_regression.detach(), proposals,
                                                        image_shapes)
            if DEFER_DETECTIONS:
                return DeferredDetections(pend), losses
            boxes, scores, labels = _resolve(pend)[0]
    elif class_logits.is_cuda:
        with torch.no_grad():
            boxes, scores, labels = postprocess_detections_concurrent(model.roi_heads, class_logits.detach(), box_regression.detach(),
                                                                      proposals, image_shapes,
                                                                      threaded=CONCURRENT_POSTPROCESS and len(proposals) > 1)
    else:
        boxes, scores, labels = model.roi_heads.postprocess_detections(class_logits, box_regression, proposals, image_shapes)
    result = [{"boxes": boxes[i], "labels": labels[i], "scores": scores[i]} for i in range(len(boxes))]
    return result, losses


def _assert_no_degenerate_boxes(targets):
    """The reference's per-image check (src/utils/eval_forward_fasterrcnn.py:40-53: one blocking host sync per image, the
    first of which waits for the whole U-Net forward).  Here the "any degenerate box" flag of the batch is computed on the
    device, copied to pinned memory without blocking, and examined at the step's next host sync (``_resolve``); the same
    assertion with the same message is raised there."""
    all_boxes = torch.cat([t["boxes"].reshape(-1, 4) for t in targets], 0)
    if all_boxes.numel() == 0:
        return

    def fail():
        for target_idx, target in enumerate(targets):
            boxes = target["boxes"]
            degenerate_boxes = boxes[:, 2:] <= boxes[:, :2]
            if degenerate_boxes.any():
                bb_idx = torch.where(degenerate_boxes.any(dim=1))[0][0]
                torch._assert(False, "All bounding boxes should have positive height and width."
                                     f" Found invalid box {boxes[bb_idx].tolist()} for target at index {target_idx}.")
    bad = (all_boxes[:, 2:] <= all_boxes[:, :2]).any()
    if not all_boxes.is_cuda:
        if bool(bad):
            fail()
        return
    flag = torch.empty((), dtype=torch.bool, pin_memory=True)
    flag.copy_(bad, non_blocking=True)
    event = torch.cuda.Event()
    event.record()
    _DEFERRED_CHECKS.append((event, flag, fail))


def _backbone_features(model, x):
    bb = model.backbone
    if STATIC_FEATURES and isinstance(bb, FrozenBackbone):
        prev = getattr(bb, "static_outputs", False)
        bb.static_outputs = True
        try:
            return bb(x)
        finally:
            bb.static_outputs = prev
    return bb(x)


def eval_forward_fasterrcnn(model, images, targets, train_det=False, model_name="fasterrcnn"):
    if not train_det and model.training:                 # (Module.eval() walks every sub-module: only when the mode changes)
        model.eval()
    _check_targets(targets)
    original_image_sizes = [tuple(img.shape[-2:]) for img in images]
    images, targets = model.transform(images, targets)
    _assert_no_degenerate_boxes(targets)
    targets_event = None
    if images.tensors.is_cuda:
        targets_event = torch.cuda.Event()
        targets_event.record()
    # the feature maps do not leave this function: the backbone may hand out its own static buffers (no 280 MB of clones)
    features = _backbone_features(model, images.tensors)
    if isinstance(features, torch.Tensor):
        features = OrderedDict([("0", features)])
    static = _static_tail_ok(model, features)
    proposals, proposal_losses = rpn_eval(model, images, features, targets, targets_event, static=static)
    detections, detector_losses = roi_heads_eval(model, features, proposals, images.image_sizes, targets)
    if isinstance(detections, (DeferredDetections, DeferredCall)):
        image_sizes = images.image_sizes
        detections.then(lambda d: model.transform.postprocess(d, image_sizes, original_image_sizes))
        if not static:
            _run_deferred_checks()
    else:
        detections = model.transform.postprocess(detections, images.image_sizes, original_image_sizes)
        _run_deferred_checks()                            # (already evaluated at the first host sync on the batched path)
    losses = {}
    losses.update(detector_losses)
    losses.update(proposal_losses)
    return losses, detections


def sigmoid_focal_loss(inputs, targets, alpha=0.25, gamma=2, reduction="none"):
    p = torch.sigmoid(inputs)
    ce_loss = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce_loss * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def compute_retinanet_loss(targets, head_outputs, anchors, model, batched=True):
    """``RetinaNet.compute_loss`` -> head ``compute_loss`` (TV models/detection/retinanet.py; reference copy
    src/utils/eval_forward_retinanet.py:64-122).  On CUDA the anchor matching runs for the whole batch (_match_batched) and
    the data-dependent counts (foreground / valid anchors per image) come back in one host read; the per-image index lists
    are then built with count-known ``nonzero`` -- same indices, same order, same reductions as the per-image loop."""
    cuda = batched and BATCHED_TAIL and anchors[0].is_cuda and all(a.shape == anchors[0].shape for a in anchors)
    if cuda and WHOLE_BATCH_RETINANET_LOSS:
        # Whole-batch form: no per-image loop and NO host sync.  Every anchor gets a (masked) loss term instead of indexing
        # the valid / foreground subsets out first, so the fp32 sums run over the same non-zero terms in a different order
        # (relative difference ~1e-6 against the per-image loop below; tests/test_heads_gpu.py).
        logits, reg = head_outputs["cls_logits"], head_outputs["bbox_regression"]            # [B, A, K], [B, A, 4]
        with torch.no_grad():
            G = max(1, max(int(t["boxes"].shape[0]) for t in targets))
            gt, present = _pad_rows([t["boxes"] for t in targets], G)
            labels_pad, _ = _pad_rows([t["labels"] for t in targets], G)
            anc = torch.stack(anchors)                                                       # [B, A, 4]
            midx = _match_batched(model.proposal_matcher, gt, present, anc)                  # [B, A]
            fg = midx >= 0
            valid = midx != model.head.classification_head.BETWEEN_THRESHOLDS
            n_fg = fg.sum(1).clamp(min=1).to(logits.dtype)
            safe = midx.clamp(min=0)
            gt_cls = torch.zeros_like(logits)
            gt_cls.scatter_(2, torch.gather(labels_pad, 1, safe)[..., None], fg[..., None].to(logits.dtype))
            matched_gt = torch.gather(gt, 1, safe[..., None].expand(-1, -1, 4))
            # (background anchors are paired with gt 0 / a zero box: their targets are never used, only kept finite)
            matched_gt = torch.where(fg[..., None], matched_gt, anc)
            target_reg = _encode_single(model.box_coder, matched_gt.reshape(-1, 4), anc.reshape(-1, 4)).reshape(anc.shape)
        cls_el = sigmoid_focal_loss(logits, gt_cls, reduction="none")
        cls = torch.where(valid[..., None], cls_el, cls_el.new_zeros(())).sum((1, 2)) / n_fg
        reg_el = F.smooth_l1_loss(reg, target_reg, reduction="none", beta=1.0)
        box = torch.where(fg[..., None], reg_el, reg_el.new_zeros(())).sum((1, 2)) / n_fg
        return {"classification": cls.sum() / len(targets), "bbox_regression": box.sum() / max(1, len(targets))}
    if cuda:
        with torch.no_grad():
            G = max(1, max(int(t["boxes"].shape[0]) for t in targets))
            gt, present = _pad_rows([t["boxes"] for t in targets], G)
            midx_all = _match_batched(model.proposal_matcher, gt, present, torch.stack(anchors))          # [B, A]
            fg_all = midx_all >= 0
            valid_all = midx_all != model.head.classification_head.BETWEEN_THRESHOLDS
            fg_count = fg_all.sum(1)
            counts = torch.cat([fg_count, valid_all.sum(1)]).tolist()                                     # the one host sync
            _run_deferred_checks()
        B = len(targets)
        matched_idxs = list(midx_all)
    else:
        matched_idxs = []
        for anchors_per_image, targets_per_image in zip(anchors, targets):
            if targets_per_image["boxes"].numel() == 0:
                matched_idxs.append(torch.full((anchors_per_image.size(0),), -1, dtype=torch.int64, device=anchors_per_image.device))
                continue
            matched_idxs.append(model.proposal_matcher(torchvision.ops.box_iou(targets_per_image["boxes"], anchors_per_image)))
    cls_losses, reg_losses = [], []
    for b, (t, logits, reg, anc, midx) in enumerate(zip(targets, head_outputs["cls_logits"], head_outputs["bbox_regression"], anchors, matched_idxs)):
        if cuda:
            n_fg, n_valid = counts[b], counts[B + b]
            fg_idx = torch.nonzero_static(fg_all[b], size=n_fg)[:, 0]
            valid_idx = torch.nonzero_static(valid_all[b], size=n_valid)[:, 0]
            gt_cls = torch.zeros_like(logits)
            if n_fg:
                gt_cls[fg_idx, t["labels"][midx[fg_idx]]] = 1.0
            # (tensor / 0-dim tensor as in the reference: ATen divides; tensor / Python scalar multiplies by the reciprocal)
            cls_losses.append(sigmoid_focal_loss(logits[valid_idx], gt_cls[valid_idx], reduction="sum") / (fg_count[b] if n_fg > 1 else 1))
            matched_gt = t["boxes"][midx[fg_idx]] if n_fg else t["boxes"].new_zeros((0, 4))
            target_regression = _encode_single(model.box_coder, matched_gt, anc[fg_idx, :])
            reg_losses.append(F.smooth_l1_loss(reg[fg_idx, :], target_regression, reduction="sum", beta=1.0) / max(1, n_fg))
            continue
        fg = midx >= 0
        num_fg = fg.sum()
        gt = torch.zeros_like(logits)
        gt[fg, t["labels"][midx[fg]]] = 1.0
        valid = midx != model.head.classification_head.BETWEEN_THRESHOLDS
        cls_losses.append(sigmoid_focal_loss(logits[valid], gt[valid], reduction="sum") / max(1, num_fg))
        fg_idx = torch.where(fg)[0]
        target_regression = model.box_coder.encode_single(t["boxes"][midx[fg_idx]], anc[fg_idx, :])
        reg_losses.append(F.smooth_l1_loss(reg[fg_idx, :], target_regression, reduction="sum", beta=1.0) / max(1, fg_idx.numel()))
    return {"classification": sum(cls_losses[1:], cls_losses[0]) / len(targets),
            "bbox_regression": sum(reg_losses[1:], reg_losses[0]) / max(1, len(targets))}


def retinanet_postprocess_detections(model, head_outputs, anchors, image_shapes):
    """``RetinaNet.postprocess_detections`` (TV models/detection/retinanet.py), same operators in the same order, with the
    box decoding on ``_decode`` (no per-call host->device scalars) and the final per-image NMS on ``batched_nms`` (hd_nms):
    identical detections."""
    from torchvision.models.detection import _utils as det_utils
    from torchvision.ops import boxes as box_ops
    class_logits, box_regression = head_outputs["cls_logits"], head_outputs["bbox_regression"]
    detections = []
    for index in range(len(image_shapes)):
        box_regression_per_image = [br[index] for br in box_regression]
        logits_per_image = [cl[index] for cl in class_logits]
        anchors_per_image, image_shape = anchors[index], image_shapes[index]
        image_boxes, image_scores, image_labels = [], [], []
        for box_regression_per_level, logits_per_level, anchors_per_level in zip(box_regression_per_image, logits_per_image, anchors_per_image):
            num_classes = logits_per_level.shape[-1]
            scores_per_level = torch.sigmoid(logits_per_level).flatten()
            keep_idxs = scores_per_level > model.score_thresh
            scores_per_level = scores_per_level[keep_idxs]
            topk_idxs = torch.where(keep_idxs)[0]
            num_topk = det_utils._topk_min(topk_idxs, model.topk_candidates, 0)
            scores_per_level, idxs = scores_per_level.topk(num_topk)
            topk_idxs = topk_idxs[idxs]
            anchor_idxs = torch.div(topk_idxs, num_classes, rounding_mode="floor")
            labels_per_level = topk_idxs % num_classes
            boxes_per_level = _decode(model.box_coder, box_regression_per_level[anchor_idxs], [anchors_per_level[anchor_idxs]])
            boxes_per_level = boxes_per_level.reshape(-1, 4)
            boxes_per_level = box_ops.clip_boxes_to_image(boxes_per_level, image_shape)
            image_boxes.append(boxes_per_level)
            image_scores.append(scores_per_level)
            image_labels.append(labels_per_level)
        image_boxes = torch.cat(image_boxes, dim=0)
        image_scores = torch.cat(image_scores, dim=0)
        image_labels = torch.cat(image_labels, dim=0)
        keep = batched_nms(image_boxes, image_scores, image_labels, model.nms_thresh)
        keep = keep[: model.detections_per_img]
        detections.append({"boxes": image_boxes[keep], "scores": image_scores[keep], "labels": image_labels[keep]})
    return detections


def retinanet_postprocess_detections_batched_begin(model, head_outputs, anchors, image_shapes):
    """``RetinaNet.postprocess_detections`` (TV models/detection/retinanet.py: per image and per level -- score threshold,
    top-k candidates, box decoding, clipping -- then per-image ``batched_nms`` and top ``detections_per_img``) for the whole
    batch: one ``topk`` per level over [B, anchors*classes] with the below-threshold scores masked out (the survivors come
    out in the same descending order as torchvision's filter-then-topk), one batched decode per level, and the batched NMS
    tail of _filter_nms_batched_begin.  ~12 launches per level and ONE host sync instead of ~14 launches and 2-3 syncs per
    (image, level) pair.  head_outputs / anchors are split per level as in eval_forward_retinanet."""
    class_logits, box_regression = head_outputs["cls_logits"], head_outputs["bbox_regression"]
    B = len(image_shapes)
    boxes_l, scores_l, labels_l, valid_l = [], [], [], []
    for lvl, (logits, reg) in enumerate(zip(class_logits, box_regression)):
        num_classes = logits.shape[-1]
        scores = torch.sigmoid(logits).flatten(1)                                   # [B, n_l * K]
        keep = scores > model.score_thresh
        k = min(model.topk_candidates, scores.shape[1])
        top_scores, top_idx = scores.masked_fill(~keep, -1.0).topk(k, dim=1)        # kept scores first, descending
        valid = top_scores > model.score_thresh
        anchor_idx = torch.div(top_idx, num_classes, rounding_mode="floor")
        labels = top_idx % num_classes
        anc = torch.stack([a[lvl] for a in anchors])                                # [B, n_l, 4]
        gi = anchor_idx[..., None].expand(-1, -1, 4)
        boxes = _decode(model.box_coder, torch.gather(reg, 1, gi).reshape(-1, 4), [torch.gather(anc, 1, gi).reshape(-1, 4)])
        boxes = _clip_boxes_batched(boxes.reshape(B, k, 4), image_shapes)
        boxes_l.append(boxes); scores_l.append(top_scores); labels_l.append(labels); valid_l.append(valid)
    boxes, scores = torch.cat(boxes_l, 1), torch.cat(scores_l, 1)
    labels, valid = torch.cat(labels_l, 1), torch.cat(valid_l, 1)
    pend = _filter_nms_batched_begin(boxes, scores, labels, valid, model.nms_thresh, model.detections_per_img)
    inner = pend.finish
    pend.finish = lambda n_sel: (lambda r: (list(r[0]), list(r[1]), list(r[2])))(inner(n_sel))
    return pend


def eval_forward_retinanet(model, images, targets, train_det=False, model_name="retinanet"):
    if not train_det and model.training:
        model.eval()
    _check_targets(targets)
    original_image_sizes = [tuple(img.shape[-2:]) for img in images]
    images, targets = model.transform(images, targets)
    _assert_no_degenerate_boxes(targets)                  # src/utils/eval_forward_retinanet.py "Check for degenerate boxes"
    features = _backbone_features(model, images.tensors)
    if isinstance(features, torch.Tensor):
        features = OrderedDict([("0", features)])
    features = list(features.values())
    head_outputs = None
    if B200_HEADS and features[0].is_cuda and isinstance(model.backbone, FrozenBackbone):
        bf16 = model.backbone.bf16_features()
        if bf16 is not None and len(bf16) == len(features) and heads.retinanet_head_towers(model.head) is not None:
            heads.USE_CUDA_GRAPH = bool(model.backbone.use_cuda_graph)
            head_outputs = heads.retinanet_head_forward(model.head, features, bf16)
    if head_outputs is None:
        head_outputs = model.head(features)
    anchors = model.anchor_generator(images, features)
    losses = compute_retinanet_loss(targets, head_outputs, anchors, model)
    num_anchors_per_level = [x.size(2) * x.size(3) for x in features]
    hw = sum(num_anchors_per_level)
    a = head_outputs["cls_logits"].size(1) // hw
    num_anchors_per_level = [n * a for n in num_anchors_per_level]
    split_head_outputs = {k: list(v.split(num_anchors_per_level, dim=1)) for k, v in head_outputs.items()}
    split_anchors = [list(x.split(num_anchors_per_level)) for x in anchors]
    n_cand = sum(min(model.topk_candidates, n * head_outputs["cls_logits"].shape[-1]) for n in num_anchors_per_level)
    if (BATCHED_TAIL and images.tensors.is_cuda and head_outputs["cls_logits"].dtype == torch.float32 and _batched_ok(n_cand)
            and all(a.shape == anchors[0].shape for a in anchors)):
        image_sizes = images.image_sizes
        det_in = {k: [t.detach() for t in v] for k, v in split_head_outputs.items()}
        with torch.no_grad():
            if DEFER_DETECTIONS and POSTPROCESS_SIDE_STREAM:
                # train step: the detections feed no loss -- they are post-processed on a side stream underneath the backward
                # pass, their data-dependent sizes read back asynchronously, the lists assembled when the trainer asks
                main = torch.cuda.current_stream(images.tensors.device)
                side = _side_streams(images.tensors.device, 1)[0]
                side.wait_stream(main)
                for t in [t for v in det_in.values() for t in v] + [a for per_image in split_anchors for a in per_image]:
                    t.record_stream(side)
                with torch.cuda.stream(side):
                    pend = retinanet_postprocess_detections_batched_begin(model, det_in, split_anchors, image_sizes)
                    deferred = DeferredDetections(pend, side)
                deferred.then(lambda d: model.transform.postprocess(d, image_sizes, original_image_sizes))
                _run_deferred_checks()
                return losses, deferred
            pend = retinanet_postprocess_detections_batched_begin(model, det_in, split_anchors, image_sizes)
            if DEFER_DETECTIONS:
                _run_deferred_checks()
                return losses, DeferredDetections(pend).then(lambda d: model.transform.postprocess(d, image_sizes, original_image_sizes))
            boxes, scores, labels = _resolve(pend)[0]
        detections = [{"boxes": boxes[i], "scores": scores[i], "labels": labels[i]} for i in range(len(boxes))]
        detections = model.transform.postprocess(detections, images.image_sizes, original_image_sizes)
        _run_deferred_checks()
        return losses, detections
    if BATCHED_TAIL and images.tensors.is_cuda:
        with torch.no_grad():
            detections = retinanet_postprocess_detections(model, split_head_outputs, split_anchors, images.image_sizes)
    else:
        detections = model.postprocess_detections(split_head_outputs, split_anchors, images.image_sizes)
    detections = model.transform.postprocess(detections, images.image_sizes, original_image_sizes)
    _run_deferred_checks()
    return losses, detections
