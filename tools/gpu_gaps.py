"""GPU idle gaps in one train step (torch.profiler kernel timeline)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200.train import HalluciDetTrainer
from hallucidet_b200 import synthetic as ostep
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = ostep.synthetic_batch(8, 512, 640, seed=123, device=dev, ir_uint8=True)
for _ in range(5):
    tr.training_step(rgb, targets, ir, targets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        out = tr.training_step(rgb, targets, ir, targets)
        float(out["total_host"])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
t0, t1 = iv[0][0], max(x[1] for x in iv)
print("span ms/step", (t1 - t0) / 3e3, "kernels/step", len(iv) / 3)
# merge busy intervals
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s, e, n in iv[1:]:
    if s > cur_e:
        gaps.append((s - cur_e, cur_e - t0, n))
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("busy ms/step", busy / 3e3, "idle ms/step", (t1 - t0 - busy) / 3e3)
big = sorted(gaps, reverse=True)[:40]
print("gaps > 100us total ms/step", sum(g for g, _, _ in gaps if g > 100) / 3e3, "count", sum(1 for g, _, _ in gaps if g > 100) / 3)
print("gaps 10..100us total ms/step", sum(g for g, _, _ in gaps if 10 < g <= 100) / 3e3, "count", sum(1 for g, _, _ in gaps if 10 < g <= 100) / 3)
print("gaps < 10us total ms/step", sum(g for g, _, _ in gaps if g <= 10) / 3e3, "count", sum(1 for g, _, _ in gaps if g <= 10) / 3)
step = (t1 - t0) / 3
for g, at, n in sorted(big, key=lambda x: x[1])[:40]:
    print(f"gap {g:8.1f} us at {at % step / 1e3:7.2f} ms into step, before {n[:80]}")

# kernel sequence of the last profiled step (name, stream, start, duration) for offline inspection
if os.environ.get("HD_GAPS_DUMP"):
    ev3 = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", -1)) for e in ev), key=lambda x: x[0])
    last = [x for x in ev3 if x[0] >= t0 + 2 * step]
    with open(os.environ["HD_GAPS_DUMP"], "w") as f:
        for s, e, n, st in last:
            f.write(f"{(s - last[0][0]):10.1f} {e - s:8.1f} s{st} {n[:140]}\n")
