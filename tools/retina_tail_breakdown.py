"""Wall-clock (synchronised) breakdown of the RetinaNet detection tail inside one config-4 train step."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200 import detection as D, heads as Hd
from hallucidet_b200.synthetic import synthetic_batch
from hallucidet_b200.train import HalluciDetTrainer
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name="retinanet", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = synthetic_batch(8, 512, 640, seed=123, device=dev)
for _ in range(4):
    tr.training_step(rgb, targets, ir, targets)
acc = {}
def wrap(mod, name):
    orig = getattr(mod, name)
    def f(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = orig(*a, **k)
        torch.cuda.synchronize(); acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        return r
    setattr(mod, name, f)
wrap(D, "compute_retinanet_loss"); wrap(D, "retinanet_postprocess_detections"); wrap(Hd, "retinanet_head_forward")
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = tr.forward_step(rgb, targets, ir, targets)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out["total"].backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    acc.setdefault("forward_step total", []).append((t1 - t0) * 1e3); acc.setdefault("backward total", []).append((t2 - t1) * 1e3)
for k, v in acc.items():
    v.sort(); print(f"{k:40s} {v[len(v)//2]:8.2f} ms")
