"""Time hd_roi_align_ml_bwd on the RoIs of a config-2 train step (HD_ROI_BWD_SEPARABLE=0/1 selects the kernel)."""
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
    rec.append((rois.detach().clone(), levels.detach().clone(), scales, [tuple(t.shape) for t in nhwc]))
    return orig(nhwc, scales, rois, levels, output_size, sampling_ratio)


ops.roi_align_ml_fwd = hook
for _ in range(3):
    tr.training_step(rgb, targets, ir, targets)
rois, levels, scales, shapes = rec[-1]
shapes_nchw = [(n, c, h, w) for (n, h, w, c) in shapes]
g = torch.randn(rois.shape[0], 256, 7, 7, device=dev)
for _ in range(3):
    ops.roi_align_ml_bwd(g, rois, levels, shapes_nchw, scales, 2, channels_last=[True] * len(shapes))
torch.cuda.synchronize()
ops.PROFILE = []
for _ in range(10):
    ops.roi_align_ml_bwd(g, rois, levels, shapes_nchw, scales, 2, channels_last=[True] * len(shapes))
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for name, _, a, b, *_ in ops.PROFILE if name == "roi_align_bwd")
print("roi_align_ml_bwd (incl. zero fill) median ms", ts[len(ts) // 2], "min", ts[0])
