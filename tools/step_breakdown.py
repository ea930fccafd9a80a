"""Wall-clock breakdown of one config-2 train step with a device synchronisation after each phase."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200.train import HalluciDetTrainer, expand_one_channel_to_output_channels  # noqa: E402
from hallucidet_b200 import detection as D  # noqa: E402
from oracle import step as ostep  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
if os.environ.get("HD_TF32", "1") == "1":
    torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = ostep.synthetic_batch(8, 512, 640, seed=123, device=dev)
for _ in range(4):
    tr.training_step(rgb, targets, ir, targets)
torch.cuda.synchronize()


def t():
    torch.cuda.synchronize()
    return time.perf_counter()


acc = {}
for it in range(5):
    tr.encoder_decoder.train()
    tr.optimizer.zero_grad(set_to_none=True)
    t0 = t()
    hal = tr.encoder_decoder(expand_one_channel_to_output_channels(ir, 3))
    t1 = t()
    model = tr.detector
    images, tg = model.transform(hal, targets)
    feats = model.backbone(images.tensors)
    t2 = t()
    proposals, pl = D.rpn_eval(model, images, feats, tg)
    t3 = t()
    det, dl = D.roi_heads_eval(model, feats, proposals, images.image_sizes, tg)
    t4 = t()
    loss = 0.1 * (pl["loss_objectness"] + pl["loss_rpn_box_reg"] + dl["loss_classifier"] + dl["loss_box_reg"])
    loss.backward()
    t5 = t()
    torch.nn.utils.clip_grad_value_(tr.encoder_decoder.parameters(), 0.5)
    tr.optimizer.step()
    t6 = t()
    for k, v in (("unet_fwd", t1 - t0), ("transform+backbone_fwd", t2 - t1), ("rpn_eval", t3 - t2), ("roi_heads_eval", t4 - t3),
                 ("backward(all)", t5 - t4), ("clip+adam", t6 - t5), ("total", t6 - t0)):
        acc.setdefault(k, []).append(v * 1e3)
for k, v in acc.items():
    v.sort()
    print(f"{k:26s} {v[len(v)//2]:8.2f} ms")
