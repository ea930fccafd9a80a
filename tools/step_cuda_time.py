"""torch.profiler view of one train step -- total GPU kernel time vs wall, top kernels, launch count."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200.train import HalluciDetTrainer
from hallucidet_b200 import synthetic as ostep
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = ostep.synthetic_batch(8, 512, 640, seed=123, device=dev)
for _ in range(4):
    tr.training_step(rgb, targets, ir, targets)
torch.cuda.synchronize()
N = 3
t0 = time.perf_counter()
for _ in range(N):
    tr.training_step(rgb, targets, ir, targets)
torch.cuda.synchronize()
print("wall ms/step (no profiler)", (time.perf_counter() - t0) / N * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        tr.training_step(rgb, targets, ir, targets)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev) if hasattr(ev[0], "device_time") else sum(e.cuda_time for e in ev)
print("cuda events/step", len(ev) / N, "total cuda ms/step", tot / N / 1e3)
agg = {}
for e in ev:
    d = e.device_time if hasattr(e, "device_time") else e.cuda_time
    a = agg.setdefault(e.name[:90], [0, 0.0]); a[0] += 1; a[1] += d
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("HD_TOP", "45"))]:
    print(f"{d / N / 1e3:8.3f} ms  n={n / N:7.1f}  {k}")
