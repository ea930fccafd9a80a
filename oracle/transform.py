"""Oracle: detector input transform (identity normalise + nearest resize to (S,S) + zero-padded batch).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  src/models/custom_generalized_transform.py:136-175  forward
  :177-186  normalize (mean [0], std [1] as configured at src/models/detector.py:43-48)
  :52-100   _resize_image_and_masks with fixed_size -> F.interpolate default mode = nearest
  :256-274  batch_images (size_divisible=1)
  :325-338  resize_boxes
"""
import math

import torch
import torch.nn.functional as F


def nearest_src_index(out_size, in_size):
    """Index map of F.interpolate(mode='nearest'): src = min(floor(dst * fp32(in/out)), in-1)."""
    scale = torch.tensor(in_size / out_size, dtype=torch.float32)
    dst = torch.arange(out_size, dtype=torch.float32)
    return torch.clamp((dst * scale).floor().to(torch.int64), max=in_size - 1)


def resize_boxes(boxes, original_size, new_size):
    """custom_generalized_transform.py:325-338 (fp32 ratios)."""
    rh = torch.tensor(new_size[0], dtype=torch.float32) / torch.tensor(original_size[0], dtype=torch.float32)
    rw = torch.tensor(new_size[1], dtype=torch.float32) / torch.tensor(original_size[1], dtype=torch.float32)
    rh, rw = rh.to(boxes.device), rw.to(boxes.device)
    xmin, ymin, xmax, ymax = boxes.unbind(1)
    return torch.stack((xmin * rw, ymin * rh, xmax * rw, ymax * rh), dim=1)


def transform_forward(images, targets=None, size=640, image_mean=(0.0,), image_std=(1.0,), size_divisible=1):
    """images: [B,3,H,W] tensor or list of [3,H,W].  Returns (batched [B,3,S,S], image_sizes, targets')."""
    out_imgs, out_tgts = [], []
    for i in range(len(images)):
        img = images[i]
        if img.dim() != 3:
            raise ValueError(f"images is expected to be a list of 3d tensors of shape [C, H, W], got {img.shape}")
        if not img.is_floating_point():
            raise TypeError(f"Expected input images to be of floating type (in range [0, 1]), but found type {img.dtype} instead")
        mean = torch.as_tensor(image_mean, dtype=img.dtype, device=img.device)
        std = torch.as_tensor(image_std, dtype=img.dtype, device=img.device)
        x = (img - mean[:, None, None]) / std[:, None, None]
        h, w = x.shape[-2:]
        x = F.interpolate(x[None], size=[size, size])[0]
        out_imgs.append(x)
        if targets is not None:
            t = dict(targets[i])
            t["boxes"] = resize_boxes(t["boxes"], (h, w), x.shape[-2:])
            out_tgts.append(t)
    image_sizes = [tuple(im.shape[-2:]) for im in out_imgs]
    mh = int(math.ceil(max(s[0] for s in image_sizes) / size_divisible) * size_divisible)
    mw = int(math.ceil(max(s[1] for s in image_sizes) / size_divisible) * size_divisible)
    batched = out_imgs[0].new_full((len(out_imgs), out_imgs[0].shape[0], mh, mw), 0)
    for i, im in enumerate(out_imgs):
        batched[i, :, : im.shape[1], : im.shape[2]].copy_(im)
    return batched, image_sizes, (out_tgts if targets is not None else None)
