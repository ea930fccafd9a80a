"""Oracle: the assembled HalluciDet train step (north-star path) in fp32 PyTorch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows train_hallucidet.py:161-209 (forward_step)
followed by ``loss.backward()`` (:283); the two extra no-grad detector passes (:183,:186) and the
plotting normalisation (:218) do not influence the loss or the gradients and are omitted.
"""
import torch

from . import unet as ounet
from . import losses as olosses
from . import detector as odet


def synthetic_batch(batch, height, width, seed=123, device="cpu"):
    """IR in [0,1), RGB in [0,1), three person boxes per image (labels 1)."""
    g = torch.Generator().manual_seed(seed)
    ir = torch.rand(batch, 1, height, width, generator=g)
    rgb = torch.rand(batch, 3, height, width, generator=g)
    sx, sy = width / 640.0, height / 512.0
    base = torch.tensor([[100., 120., 180., 300.], [400., 200., 450., 330.], [20., 30., 60., 140.]])
    boxes = base * torch.tensor([sx, sy, sx, sy])
    targets = [{"boxes": boxes.clone().to(device), "labels": torch.ones(3, dtype=torch.int64, device=device)} for _ in range(batch)]
    return ir.to(device), rgb.to(device), targets


def train_step(unet_state, detector, ir, rgb, targets, size=640, detector_name="fasterrcnn",
               pixel=None, weights=None, det_seed=7, training=True, backbone_fn=None, unet_fn=None):
    """One forward+backward.  Returns dict(loss, parts, hal, grads{param_key: grad}).

    unet_state: oracle/unet.py state dict (parameter tensors get requires_grad here).
    """
    weights = dict(olosses.DEFAULT_WEIGHTS, **(weights or {}))
    params = {k: v for k, v in unet_state.items() if ounet.is_param(k)}
    for v in params.values():
        v.requires_grad_(True)
        v.grad = None
    ir3 = ounet.expand_ir(ir, 3)                                            # train_hallucidet.py:170
    hal = (unet_fn or (lambda x: ounet.unet_forward(unet_state, x, training=training)))(ir3)   # :171
    hal.retain_grad()
    l_rgb, l_ir = olosses.regulariser(pixel, rgb, ir3, hal, weights["pixel_rgb"], weights["pixel_ir"])   # :173-176
    torch.manual_seed(det_seed)
    losses_det, _ = odet.calculate_loss(detector, hal, targets, size, detector_name, backbone_fn)      # :180
    det_total, parts = olosses.assemble_detection_loss(losses_det, detector_name, weights)             # :189-205
    total = det_total + l_rgb + l_ir                                                                   # :209
    total.backward()
    grads = {k: v.grad.detach().clone() for k, v in params.items() if v.grad is not None}
    return {"loss": total.detach(), "parts": {k: (v.detach() if torch.is_tensor(v) else v) for k, v in parts.items()},
            "pixel_rgb": l_rgb, "pixel_ir": l_ir, "hal": hal.detach(), "dhal": hal.grad.detach().clone(), "grads": grads}
