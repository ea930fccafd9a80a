"""Synthetic HalluciDet batches: the shapes and value ranges the reference's dataloader delivers
(src/dataloader/dataloader.py:13-73: IR as one uint8 plane scaled by 1/255 -> fp32 [B,1,H,W] in [0,1]; RGB fp32 [B,3,H,W]
in [0,1]; per image a dict of xyxy pixel boxes and int64 labels == 1, src/utils/utils.py:433-436).

There is no dataset on the benchmark box, so bench.py / smoke() / the tests draw these from a seeded generator.
"""
import torch


def synthetic_batch(batch, height, width, seed=123, device="cpu", ir_uint8=False):
    """IR in [0,1) (or the uint8 plane it is decoded from when ``ir_uint8``), RGB in [0,1), three person boxes per image."""
    g = torch.Generator().manual_seed(seed)
    ir = torch.rand(batch, 1, height, width, generator=g)
    rgb = torch.rand(batch, 3, height, width, generator=g)
    if ir_uint8:
        ir = (ir * 255.0).round().to(torch.uint8)          # what the camera frame holds; /255 happens on the device
    sx, sy = width / 640.0, height / 512.0
    base = torch.tensor([[100., 120., 180., 300.], [400., 200., 450., 330.], [20., 30., 60., 140.]])
    boxes = base * torch.tensor([sx, sy, sx, sy])
    targets = [{"boxes": boxes.clone().to(device), "labels": torch.ones(3, dtype=torch.int64, device=device)} for _ in range(batch)]
    return ir.to(device), rgb.to(device), targets
