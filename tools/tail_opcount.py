"""kernels launched and CPU time per section of the detection tail."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200.train import HalluciDetTrainer, expand_one_channel_to_output_channels
from hallucidet_b200 import detection as D
from oracle import step as ostep
from torch.profiler import profile, ProfilerActivity, record_function
from torchvision.models.detection.rpn import concat_box_prediction_layers
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = ostep.synthetic_batch(8, 512, 640, seed=123, device=dev)
for _ in range(4):
    tr.training_step(rgb, targets, ir, targets)
model = tr.detector
def run():
    tr.encoder_decoder.train()
    hal = tr.encoder_decoder(expand_one_channel_to_output_channels(ir, 3))
    images, tg = model.transform(hal, targets)
    feats = model.backbone(images.tensors)
    fl = list(feats.values())
    with record_function("S:rpn.head"): obj, deltas = model.rpn.head(fl)
    with record_function("S:anchors"): anchors = model.rpn.anchor_generator(images, fl)
    napl = [o[0].shape[0] * o[0].shape[1] * o[0].shape[2] for o in obj]
    with record_function("S:concat"): o2, d2 = concat_box_prediction_layers(obj, deltas)
    with record_function("S:decode"): props = model.rpn.box_coder.decode(d2.detach(), anchors).view(len(anchors), -1, 4)
    with torch.no_grad():
        with record_function("S:assign"): labels, mgt = D.assign_targets_to_anchors_batched(model.rpn, anchors, tg)
        with record_function("S:encode"): rt = model.rpn.box_coder.encode_single(mgt.reshape(-1, 4), torch.cat(anchors, 0))
        with record_function("S:sample_begin"): ps = D._sample_batched_begin(model.rpn.fg_bg_sampler, labels)
    with record_function("S:filter_begin"): pf = D.filter_proposals_batched_begin(model.rpn, props, o2, images.image_sizes, napl)
    with record_function("S:resolve_rpn"): (boxes, scores), samples = D._resolve(pf, ps)
    with record_function("S:rpn_loss"): lo, lr = D.rpn_compute_loss_batched(model.rpn, o2, d2, labels, rt, samples=samples)
    with torch.no_grad():
        with record_function("S:select_training"): p2, midx, lab, regt, npos = D.select_training_samples_batched(model.roi_heads, boxes, tg, return_num_pos=True)
    with record_function("S:roi_pool"): bf = D.multiscale_roi_align_one_sync(model.roi_heads.box_roi_pool, feats, p2, images.image_sizes)
    with record_function("S:box_head+pred"): cl, br = model.roi_heads.box_predictor(model.roi_heads.box_head(bf))
    with record_function("S:fastrcnn_loss"): lc, lb = D.fastrcnn_loss_static(cl, br, lab, regt, npos)
    with torch.no_grad():
        with record_function("S:postprocess_begin"): pend = D.postprocess_detections_batched_begin(model.roi_heads, cl.detach(), br.detach(), p2, images.image_sizes)
    with record_function("S:backward"): (0.1 * (lo + lr + lc + lb)).backward()
    with record_function("S:post_resolve"): D._resolve(pend)
    tr.optimizer.zero_grad(set_to_none=True)
run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
evs = prof.events()
secs = [e for e in evs if e.name.startswith("S:")]
launches = [e for e in evs if e.name in ("cudaLaunchKernel", "cudaMemcpyAsync", "cudaLaunchKernelExC", "cudaMemsetAsync", "cudaGraphLaunch")]
for s in secs:
    n = sum(1 for l in launches if s.time_range.start <= l.time_range.start <= s.time_range.end)
    print(f"{s.name:24s} cpu {s.cpu_time_total / 1e3:7.2f} ms  launches {n}")
want = os.environ.get("HD_SECTION")
if want:
    from collections import Counter
    sec = [s for s in secs if s.name == "S:" + want][0]
    names = Counter()
    for e in evs:
        if e.device_type == torch.autograd.DeviceType.CUDA and False:
            pass
    kern = [e for e in evs if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::")
            and sec.time_range.start <= e.time_range.start <= sec.time_range.end]
    # top-level aten ops only (not nested inside another aten op of the section)
    kern.sort(key=lambda e: e.time_range.start)
    top, end = [], -1
    for e in kern:
        if e.time_range.start >= end:
            top.append(e)
            end = e.time_range.end
    for e in top:
        n = sum(1 for l in launches if e.time_range.start <= l.time_range.start <= e.time_range.end)
        names[(e.name, n)] += 1
    for (nm, n), c in sorted(names.items(), key=lambda kv: -kv[1] * max(kv[0][1], 1)):
        print(f"  {nm:40s} launches/op {n}  x{c}")
