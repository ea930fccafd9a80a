"""GPU-time breakdown of the config-2 train step by phase, WITHOUT synchronising inside the step: CUDA events are recorded
on the main stream around the four kernel programs (U-Net forward / backward, backbone forward / backward; CUDA-graph
replays) and the optimizer; the gaps between them are the detection tail (forward and backward) and the transform.
Usage: python tools/step_phases.py [detector] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200 import backbone as hb, unet as hu  # noqa: E402
from hallucidet_b200.synthetic import synthetic_batch  # noqa: E402
from hallucidet_b200.train import HalluciDetTrainer  # noqa: E402

detector = sys.argv[1] if len(sys.argv) > 1 else "fasterrcnn"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name=detector, size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = synthetic_batch(8, 512, 640, seed=123, device=dev)
for _ in range(5):
    tr.training_step(rgb, targets, ir, targets)
torch.cuda.synchronize()

marks = []


def wrap(cls, name, tag):
    orig = getattr(cls, name)

    def f(self, *a, **k):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(self, *a, **k)
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        marks.append((tag, e0, e1))
        return out

    setattr(cls, name, f)


wrap(hu._UnetEngine, "forward", "unet_fwd")
wrap(hu._UnetEngine, "backward", "unet_bwd")
wrap(hb._BackboneEngine, "forward", "backbone_fwd")
wrap(hb._BackboneEngine, "backward", "backbone_bwd")
opt_step = tr.optimizer.step


def timed_opt(*a, **k):
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    r = opt_step(*a, **k)
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    marks.append(("adam", e0, e1))
    return r


tr.optimizer.step = timed_opt
acc = {}
for _ in range(steps):
    marks.clear()
    s0 = torch.cuda.Event(enable_timing=True)
    s0.record()
    tr.training_step(rgb, targets, ir, targets)
    s1 = torch.cuda.Event(enable_timing=True)
    s1.record()
    torch.cuda.synchronize()
    d = {t: a.elapsed_time(b) for t, a, b in marks}
    ev = {t: (a, b) for t, a, b in marks}
    d["tail_fwd (transform, rpn, roi, losses)"] = ev["backbone_fwd"][1].elapsed_time(ev["backbone_bwd"][0]) if "backbone_bwd" in ev else 0.0
    d["unet_fwd -> backbone_fwd (transform)"] = ev["unet_fwd"][1].elapsed_time(ev["backbone_fwd"][0])
    d["backbone_bwd -> unet_bwd (transform bwd)"] = ev["backbone_bwd"][1].elapsed_time(ev["unet_bwd"][0])
    d["unet_bwd -> adam (allreduce, clip)"] = ev["unet_bwd"][1].elapsed_time(ev["adam"][0])
    d["step"] = s0.elapsed_time(s1)
    for k, v in d.items():
        acc.setdefault(k, []).append(v)
print(f"phase GPU time on the main stream, median of {steps} steps (detector {detector}); tail_fwd includes the tail's backward up to the backbone:")
for k, v in acc.items():
    v.sort()
    print(f"  {k:44s} {v[len(v) // 2]:8.3f} ms")
