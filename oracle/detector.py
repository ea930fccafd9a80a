"""Oracle: frozen-detector loss in eval mode (Faster R-CNN / RetinaNet over torchvision heads).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  src/models/detector.py:25-66,122-141            detector construction (2-class re-heading, Xavier re-init)
  src/utils/eval_forward_fasterrcnn.py:13-68      eval_forward_fasterrcnn; :72-102 rpn_eval; :105-185 roi_heads_eval
  src/utils/eval_forward_retinanet.py:22-50       sigmoid_focal_loss; :83-160 eval_forward_retinanet; :163-244 losses
The RPN / RoI / RetinaNet heads, anchor generator, matcher, samplers and box coder are torchvision's own
objects (third-party, the reference calls them unchanged); the backbone and the transform are the oracle's
functional restatements (oracle/backbone.py, oracle/transform.py) unless ``backbone_fn`` is given.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F
import torchvision
from torchvision.models.detection.image_list import ImageList
from torchvision.models.detection.roi_heads import fastrcnn_loss
from torchvision.models.detection.rpn import concat_box_prediction_layers

from . import backbone as obb
from . import transform as otr


def build_detector(name="fasterrcnn", seed=123, n_classes=2):
    """src/models/detector.py:25-66 with random-init weights (no network): returns the torchvision module."""
    torch.manual_seed(seed)
    if "fasterrcnn" in name:
        det = torchvision.models.detection.fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None)
        in_features = det.roi_heads.box_predictor.cls_score.in_features
        det.roi_heads.box_predictor = torchvision.models.detection.faster_rcnn.FastRCNNPredictor(in_features, n_classes)
        for layer in det.roi_heads.modules():                      # detector.py:15-20 (_xavier_init)
            if isinstance(layer, torch.nn.Conv2d):
                torch.nn.init.xavier_uniform_(layer.weight)
                if layer.bias is not None:
                    torch.nn.init.constant_(layer.bias, 0.0)
    elif "retinanet" in name:
        det = torchvision.models.detection.retinanet_resnet50_fpn(weights=None, weights_backbone=None)
        out_channels = det.head.classification_head.conv[0].out_channels
        num_anchors = det.head.classification_head.num_anchors
        det.head.classification_head.num_classes = n_classes
        cls_logits = torch.nn.Conv2d(out_channels, num_anchors * n_classes, kernel_size=3, stride=1, padding=1)
        torch.nn.init.normal_(cls_logits.weight, std=0.01)
        torch.nn.init.constant_(cls_logits.bias, -math.log((1 - 0.01) / 0.01))
        det.head.classification_head.cls_logits = cls_logits
    else:
        raise ValueError(name)
    det.eval()
    for p in det.parameters():
        p.requires_grad_(False)
    return det


def randomize_bn_stats(det, seed=7):
    """Give the frozen BN layers non-trivial statistics/affines (random init leaves them at 0/1)."""
    g = torch.Generator().manual_seed(seed)
    for m in det.backbone.modules():
        if hasattr(m, "running_mean") and m.running_mean is not None:
            n = m.running_mean.numel()
            m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=g) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(n, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(n, generator=g) * 0.1)


def _oracle_backbone(det, variant):
    state = det.backbone.state_dict()
    return lambda x: obb.backbone_forward(state, x, variant=variant)


def rpn_eval(model, images, features, targets):
    """eval_forward_fasterrcnn.py:72-102."""
    features = list(features.values())
    objectness, pred_bbox_deltas = model.rpn.head(features)
    anchors = model.rpn.anchor_generator(images, features)
    num_images = len(anchors)
    shapes = [o[0].shape for o in objectness]
    num_anchors_per_level = [s[0] * s[1] * s[2] for s in shapes]
    objectness, pred_bbox_deltas = concat_box_prediction_layers(objectness, pred_bbox_deltas)
    proposals = model.rpn.box_coder.decode(pred_bbox_deltas.detach(), anchors)
    proposals = proposals.view(num_images, -1, 4)
    boxes, _ = model.rpn.filter_proposals(proposals, objectness, images.image_sizes, num_anchors_per_level)
    labels, matched_gt_boxes = model.rpn.assign_targets_to_anchors(anchors, targets)
    regression_targets = model.rpn.box_coder.encode(matched_gt_boxes, anchors)
    loss_objectness, loss_rpn_box_reg = model.rpn.compute_loss(objectness, pred_bbox_deltas, labels, regression_targets)
    return boxes, {"loss_objectness": loss_objectness, "loss_rpn_box_reg": loss_rpn_box_reg}


def roi_heads_eval(model, features, proposals, image_shapes, targets):
    """eval_forward_fasterrcnn.py:105-147 (box branch; no keypoint head on this detector)."""
    proposals, matched_idxs, labels, regression_targets = model.roi_heads.select_training_samples(proposals, targets)
    box_features = model.roi_heads.box_roi_pool(features, proposals, image_shapes)
    box_features = model.roi_heads.box_head(box_features)
    class_logits, box_regression = model.roi_heads.box_predictor(box_features)
    loss_classifier, loss_box_reg = fastrcnn_loss(class_logits, box_regression, labels, regression_targets)
    boxes, scores, labels = model.roi_heads.postprocess_detections(class_logits, box_regression, proposals, image_shapes)
    result = [{"boxes": boxes[i], "labels": labels[i], "scores": scores[i]} for i in range(len(boxes))]
    return result, {"loss_classifier": loss_classifier, "loss_box_reg": loss_box_reg}


def _postprocess(result, image_shapes, original_image_sizes):
    """custom_generalized_transform.py:276-299 (eval mode)."""
    for i, (pred, im_s, o_im_s) in enumerate(zip(result, image_shapes, original_image_sizes)):
        result[i]["boxes"] = otr.resize_boxes(pred["boxes"], im_s, o_im_s)
    return result


def eval_forward_fasterrcnn(model, images, targets, size, backbone_fn=None):
    """eval_forward_fasterrcnn.py:13-68.  images [B,3,H,W]; returns (losses, detections)."""
    model.eval()
    original_image_sizes = [tuple(img.shape[-2:]) for img in images]
    batched, image_sizes, targets = otr.transform_forward(images, targets, size=size)
    for target_idx, target in enumerate(targets):
        boxes = target["boxes"]
        if (boxes[:, 2:] <= boxes[:, :2]).any():
            raise AssertionError(f"All bounding boxes should have positive height and width (target {target_idx}).")
    image_list = ImageList(batched, image_sizes)
    backbone_fn = backbone_fn or _oracle_backbone(model, "fasterrcnn")
    features = backbone_fn(batched)
    proposals, proposal_losses = rpn_eval(model, image_list, features, targets)
    detections, detector_losses = roi_heads_eval(model, features, proposals, image_list.image_sizes, targets)
    detections = _postprocess(detections, image_list.image_sizes, original_image_sizes)
    losses = {}
    losses.update(detector_losses)
    losses.update(proposal_losses)
    return losses, detections


def sigmoid_focal_loss(inputs, targets, alpha=0.25, gamma=2):
    """eval_forward_retinanet.py:22-50 with reduction='sum'."""
    p = torch.sigmoid(inputs)
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    alpha_t = alpha * targets + (1 - alpha) * (1 - targets)
    return (alpha_t * loss).sum()


def retinanet_losses(model, targets, head_outputs, anchors):
    """eval_forward_retinanet.py:163-244."""
    matched_idxs = []
    for anchors_per_image, t in zip(anchors, targets):
        if t["boxes"].numel() == 0:
            matched_idxs.append(torch.full((anchors_per_image.size(0),), -1, dtype=torch.int64, device=anchors_per_image.device))
            continue
        matched_idxs.append(model.proposal_matcher(torchvision.ops.box_iou(t["boxes"], anchors_per_image)))
    cls_losses, reg_losses = [], []
    for t, logits, reg, anc, midx in zip(targets, head_outputs["cls_logits"], head_outputs["bbox_regression"], anchors, matched_idxs):
        fg = midx >= 0
        num_fg = fg.sum()
        gt = torch.zeros_like(logits)
        gt[fg, t["labels"][midx[fg]]] = 1.0
        valid = midx != model.head.classification_head.BETWEEN_THRESHOLDS
        cls_losses.append(sigmoid_focal_loss(logits[valid], gt[valid]) / max(1, num_fg))
        fg_idx = torch.where(fg)[0]
        tgt = model.box_coder.encode_single(t["boxes"][midx[fg_idx]], anc[fg_idx, :])
        reg_losses.append(F.smooth_l1_loss(reg[fg_idx, :], tgt, reduction="sum", beta=1.0) / max(1, fg_idx.numel()))
    return {"classification": sum(cls_losses[1:], cls_losses[0]) / len(targets),
            "bbox_regression": sum(reg_losses[1:], reg_losses[0]) / max(1, len(targets))}


def eval_forward_retinanet(model, images, targets, size, backbone_fn=None):
    """eval_forward_retinanet.py:83-160."""
    model.eval()
    original_image_sizes = [tuple(img.shape[-2:]) for img in images]
    batched, image_sizes, targets = otr.transform_forward(images, targets, size=size)
    image_list = ImageList(batched, image_sizes)
    backbone_fn = backbone_fn or _oracle_backbone(model, "retinanet")
    features = list(backbone_fn(batched).values())
    head_outputs = model.head(features)
    anchors = model.anchor_generator(image_list, features)
    losses = retinanet_losses(model, targets, head_outputs, anchors)
    num_anchors_per_level = [x.size(2) * x.size(3) for x in features]
    hw = sum(num_anchors_per_level)
    a = head_outputs["cls_logits"].size(1) // hw
    num_anchors_per_level = [n * a for n in num_anchors_per_level]
    split_head_outputs = {k: list(v.split(num_anchors_per_level, dim=1)) for k, v in head_outputs.items()}
    split_anchors = [list(x.split(num_anchors_per_level)) for x in anchors]
    detections = model.postprocess_detections(split_head_outputs, split_anchors, image_list.image_sizes)
    detections = _postprocess(detections, image_list.image_sizes, original_image_sizes)
    return losses, detections


def calculate_loss(model, images, targets, size, model_name="fasterrcnn", backbone_fn=None):
    """src/models/detector.py:104-118."""
    if "fasterrcnn" in model_name:
        return eval_forward_fasterrcnn(model, images, targets, size, backbone_fn)
    if "retinanet" in model_name:
        return eval_forward_retinanet(model, images, targets, size, backbone_fn)
    raise ValueError(model_name)
