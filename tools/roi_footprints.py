"""Footprint (feature pixels touched) of the RoIs the box head pools in a config-2 train step: decides how many of them a
shared-memory accumulation tile can serve in roi_align backward."""
import sys, torch
sys.path.insert(0, ".")
from hallucidet_b200 import ops
from hallucidet_b200.train import HalluciDetTrainer
from hallucidet_b200.synthetic import synthetic_batch

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = synthetic_batch(8, 512, 640, seed=123, device=dev, ir_uint8=True)
rec = []
orig = ops.roi_align_ml_fwd


def hook(nhwc, scales, rois, levels, output_size, sampling_ratio):
    rec.append((rois.detach().cpu(), levels.detach().cpu(), scales))
    return orig(nhwc, scales, rois, levels, output_size, sampling_ratio)


ops.roi_align_ml_fwd = hook
for step in range(12):
    tr.training_step(rgb, targets, ir, targets)
    if step in (0, 5, 11):
        rois, levels, scales = rec[-1]
        sc = torch.tensor(scales)[levels]
        w = (rois[:, 3] - rois[:, 1]) * sc
        h = (rois[:, 4] - rois[:, 2]) * sc
        fw, fh = w.clamp(min=1).floor() + 2, h.clamp(min=1).floor() + 2
        F = fw * fh
        print(f"step {step}: {len(rois)} rois, levels {torch.bincount(levels, minlength=4).tolist()}, footprint px: median {F.median():.0f} "
              f"mean {F.mean():.0f} | <=64: {(F <= 64).float().mean():.2f} <=128: {(F <= 128).float().mean():.2f} "
              f"<=168: {(F <= 168).float().mean():.2f} <=256: {(F <= 256).float().mean():.2f} <=400: {(F <= 400).float().mean():.2f} <=900: {(F <= 900).float().mean():.2f}")
