"""Oracle: pixel regulariser and loss assembly.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  src/losses/losses.py:28-48                nn.MSELoss / nn.L1Loss, mean reduction
  train_hallucidet.py:173-176               loss_pixel(reference_image, hallucinated) * weight
  train_hallucidet.py:189-209               detection-loss weighting and total
  src/config/config.py:58-69                default weights (0.1 x4 detection, 0 regulariser)
"""
import torch
import torch.nn.functional as F

DEFAULT_WEIGHTS = {
    "pixel_rgb": 0.0, "pixel_ir": 0.0, "perceptual_rgb": 0.0, "perceptual_ir": 0.0,
    "det_regression": 0.1, "det_classification": 0.1, "det_objectness": 0.1,
    "det_rpn_box_reg": 0.1, "det_bbox_ctrness": 0.1,
}


def pixel_loss(kind, ref, hal):
    if kind == "mse":
        return F.mse_loss(ref, hal)
    if kind == "l1":
        return F.l1_loss(ref, hal)
    raise ValueError(kind)


def regulariser(kind, rgb, ir3, hal, w_rgb, w_ir):
    """train_hallucidet.py:173,175: (pixel(rgb,hal)*w_rgb, pixel(ir3,hal)*w_ir); kind None -> (0.0, 0.0)."""
    if kind is None:
        return 0.0, 0.0
    return pixel_loss(kind, rgb, hal) * w_rgb, pixel_loss(kind, ir3, hal) * w_ir


def assemble_detection_loss(losses_det, detector_name, weights=DEFAULT_WEIGHTS):
    """train_hallucidet.py:189-205."""
    d = dict(losses_det)
    if "fasterrcnn" in detector_name:
        d["classification"] = d["loss_classifier"]
        d["bbox_regression"] = d["loss_box_reg"]
    d["bbox_regression"] = d["bbox_regression"] * weights["det_regression"]
    d["classification"] = d["classification"] * weights["det_classification"]
    d["loss_objectness"] = d["loss_objectness"] * weights["det_objectness"] if "fasterrcnn" in detector_name else 0.0
    d["loss_rpn_box_reg"] = d["loss_rpn_box_reg"] * weights["det_rpn_box_reg"] if "fasterrcnn" in detector_name else 0.0
    d["bbox_ctrness"] = d["bbox_ctrness"] * weights["det_bbox_ctrness"] if "fcos" in detector_name else 0.0
    total = d["bbox_regression"] + d["classification"] + d["loss_objectness"] + d["loss_rpn_box_reg"] + d["bbox_ctrness"]
    return total, d
