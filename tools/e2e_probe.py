"""Where does the end-to-end leg of bench.py lose time against the device-resident leg?  (config 4 showed 26 vs 15.6 ms)
usage: python tools/e2e_probe.py [fasterrcnn|retinanet]"""
import sys, time, torch
sys.path.insert(0, ".")
from hallucidet_b200.train import HalluciDetTrainer, DevicePrefetcher
from hallucidet_b200.synthetic import synthetic_batch

det = sys.argv[1] if len(sys.argv) > 1 else "retinanet"
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name=det, size=640, seed=123, device=dev, use_cuda_graph=True)
ir_h, rgb_h, targets = synthetic_batch(8, 512, 640, seed=123, ir_uint8=True)
ir_h, rgb_h = ir_h.pin_memory(), rgb_h.pin_memory()
targets = [{k: v.to(dev) for k, v in t.items()} for t in targets]
ir_d, rgb_d = ir_h.to(dev), rgb_h.to(dev)
pf = DevicePrefetcher(dev)


def resident():
    return tr.training_step(rgb_d, targets, ir_d, targets)


def resident_read():
    float(tr.training_step(rgb_d, targets, ir_d, targets)["total_host"])


def prefetch_only():
    if pf.pending is None:
        pf.put(ir_h, rgb_h)
    ir, rgb = pf.get()
    pf.put(ir_h, rgb_h)
    tr.training_step(rgb, targets, ir, targets)


def full():
    if pf.pending is None:
        pf.put(ir_h, rgb_h)
    ir, rgb = pf.get()
    pf.put(ir_h, rgb_h)
    float(tr.training_step(rgb, targets, ir, targets)["total_host"])


def copy_same_stream():
    ir, rgb = ir_h.to(dev, non_blocking=True), rgb_h.to(dev, non_blocking=True)
    float(tr.training_step(rgb, targets, ir, targets)["total_host"])


for _ in range(6):
    resident()
for name, fn in [("resident", resident), ("resident+read", resident_read), ("prefetch", prefetch_only), ("prefetch+read", full),
                 ("same-stream copy+read", copy_same_stream), ("resident", resident)]:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    print(f"{det} {name:24s} {(time.perf_counter() - t0) / 20 * 1e3:7.2f} ms/step", flush=True)
