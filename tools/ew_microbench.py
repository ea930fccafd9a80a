"""Element-wise kernels at config-2 shapes: max-pool backward (arg-max form), 2x up-sampling, timed back to back in a CUDA graph."""
import sys, torch
sys.path.insert(0, ".")
from hallucidet_b200 import ops
dev = torch.device("cuda", 0)


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    return sorted(ts)[len(ts) // 2]


for name, (n, h, w, c), use_add in (("unet stem 256x320x64 +skip grad", (8, 256, 320, 64), True), ("backbone stem 320x320x64", (8, 320, 320, 64), False)):
    x, y = bf(n, h, w, c), bf(n, h // 2, w // 2, c)
    idx = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=dev)
    ops.maxpool_fwd(x, y, idx=idx, mask_nonpositive=True)
    dy, dx = bf(n, h // 2, w // 2, c), bf(n, h, w, c)
    add = bf(n, h, w, c) if use_add else None
    us = timed(lambda: ops.maxpool_bwd(x, y, dy, dx, add=add, idx=idx))
    nbytes = 2 * (dx.numel() * (2 if use_add else 1) + dy.numel()) + idx.numel()
    print(f"maxpool_bwd {name:34s} {us:7.1f} us  {nbytes / us / 1e3:7.0f} GB/s (algorithmic)")
for name, (n, h, w, c) in (("unet", (8, 256, 320, 64)), ("backbone", (8, 320, 320, 64))):
    x, y = bf(n, h, w, c), bf(n, h // 2, w // 2, c)
    idx = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=dev)
    us = timed(lambda: ops.maxpool_fwd(x, y, idx=idx, mask_nonpositive=True))
    print(f"maxpool_fwd {name:10s} {us:7.1f} us  {(2 * (x.numel() + y.numel()) + idx.numel()) / us / 1e3:7.0f} GB/s")
xr, yr = torch.rand(8, 3, 512, 640, device=dev), torch.empty(8, 3, 640, 640, device=dev)
mean, std = torch.tensor([0.485, 0.456, 0.406], device=dev), torch.tensor([0.229, 0.224, 0.225], device=dev)
us = timed(lambda: ops.resize_nearest_fwd(xr, yr, mean, std))
print(f"resize_fwd 512x640 -> 640x640 {us:7.1f} us")
gx = torch.empty_like(xr)
us = timed(lambda: ops.resize_nearest_bwd(yr, gx, std))
print(f"resize_bwd 640x640 -> 512x640 {us:7.1f} us")
for (n, h, w, c) in ((8, 256, 320, 32), (8, 128, 160, 64), (8, 64, 80, 128), (8, 32, 40, 256), (8, 16, 20, 512)):
    x, y = bf(n, h, w, c), bf(n, 2 * h, 2 * w, c)
    us = timed(lambda: ops.upsample2x_fwd(x, y))
    print(f"upsample2x_fwd [{n},{h},{w},{c}] {us:7.1f} us  {2 * (x.numel() + y.numel()) / us / 1e3:7.0f} GB/s")
    us = timed(lambda: ops.upsample2x_bwd(y, x))
    print(f"upsample2x_bwd [{n},{h},{w},{c}] {us:7.1f} us  {2 * (x.numel() + y.numel()) / us / 1e3:7.0f} GB/s")
