"""Sync-timed breakdown of the torchvision detection tail (rpn_eval / roi_heads_eval) at config-2 shapes."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200.train import HalluciDetTrainer, expand_one_channel_to_output_channels  # noqa
from hallucidet_b200 import detection as D  # noqa
from hallucidet_b200 import synthetic as ostep  # noqa
from torchvision.models.detection.rpn import concat_box_prediction_layers
from torchvision.models.detection.roi_heads import fastrcnn_loss

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = ostep.synthetic_batch(8, 512, 640, seed=123, device=dev)
for _ in range(3):
    tr.training_step(rgb, targets, ir, targets)


def t():
    torch.cuda.synchronize()
    return time.perf_counter()


acc = {}
def rec(k, dt):
    acc.setdefault(k, []).append(dt * 1e3)

model = tr.detector
for it in range(4):
    tr.encoder_decoder.train()
    hal = tr.encoder_decoder(expand_one_channel_to_output_channels(ir, 3))
    images, tg = model.transform(hal, targets)
    feats = model.backbone(images.tensors)
    fl = list(feats.values())
    t0 = t(); obj, deltas = model.rpn.head(fl); t1 = t(); rec("rpn.head", t1 - t0)
    anchors = model.rpn.anchor_generator(images, fl); t2 = t(); rec("anchor_generator", t2 - t1)
    napl = [o[0].shape[0] * o[0].shape[1] * o[0].shape[2] for o in obj]
    o2, d2 = concat_box_prediction_layers(obj, deltas); t3 = t(); rec("concat_box_prediction_layers", t3 - t2)
    props = model.rpn.box_coder.decode(d2.detach(), anchors).view(len(anchors), -1, 4); t4 = t(); rec("decode", t4 - t3)
    boxes, scores = D.filter_proposals_batched(model.rpn, props, o2, images.image_sizes, napl); t5 = t(); rec("filter_proposals_batched", t5 - t4)
    with torch.no_grad():
        labels, mgt = D.assign_targets_to_anchors_batched(model.rpn, anchors, tg); t6 = t(); rec("assign_targets_to_anchors_batched", t6 - t5)
        rt = model.rpn.box_coder.encode_single(mgt.reshape(-1, 4), torch.cat(anchors, 0)); t7 = t(); rec("encode", t7 - t6)
    lo, lr = D.rpn_compute_loss_batched(model.rpn, o2, d2, labels, rt); t8 = t(); rec("rpn_compute_loss_batched", t8 - t7)
    with torch.no_grad():
        p2, midx, lab, regt = D.select_training_samples_batched(model.roi_heads, boxes, tg); t9 = t(); rec("select_training_samples_batched", t9 - t8)
    bf = model.roi_heads.box_roi_pool(feats, p2, images.image_sizes); t10 = t(); rec("box_roi_pool", t10 - t9)
    bf = model.roi_heads.box_head(bf); t11 = t(); rec("box_head", t11 - t10)
    cl, br = model.roi_heads.box_predictor(bf); t12 = t(); rec("box_predictor", t12 - t11)
    lc, lb = fastrcnn_loss(cl, br, lab, regt); t13 = t(); rec("fastrcnn_loss", t13 - t12)
    with torch.no_grad():
        D.postprocess_detections_batched(model.roi_heads, cl.detach(), br.detach(), p2, images.image_sizes)
    t14 = t(); rec("postprocess_batched", t14 - t13)
    (0.1 * (lo + lr + lc + lb)).backward(); t15 = t(); rec("backward(all)", t15 - t14)
    tr.optimizer.zero_grad(set_to_none=True)
for k, v in acc.items():
    v.sort(); print(f"{k:34s} {v[len(v)//2]:8.2f} ms")
