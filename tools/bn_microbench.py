"""Micro-benchmark of the BatchNorm element-wise kernels at the largest config-2 shapes (CUDA events, L2 flush)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200 import ops
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bf(*s): return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
def train(fn, n=20):
    """n back-to-back launches captured in a CUDA graph and replayed (no host launch cost; PDL-chained; operands L2-warm
    as in the step, where the producer has just written them): GPU us per launch."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    ts.sort(); return ts[len(ts) // 2]
barrier = torch.zeros(2, dtype=torch.int32, device=dev)
for (n, h, w, c) in [(8, 512, 640, 16), (8, 256, 320, 32), (8, 256, 320, 64), (8, 128, 160, 64), (8, 64, 80, 128), (8, 32, 40, 256), (8, 16, 20, 512)]:
    z, g, y, dz = bf(n, h, w, c), bf(n, h, w, c), bf(n, h, w, c), bf(n, h, w, c)
    mean, invstd = torch.randn(c, device=dev) * 0.1, torch.rand(c, device=dev) + 0.5
    scale, shift, gamma = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1, torch.rand(c, device=dev) + 0.5
    sums = torch.zeros(2, c, device=dev)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    nbytes = z.numel() * 2
    t_apply = timeit(lambda: ops.bn_apply(z, scale, shift, y, relu=True))
    t_red = timeit(lambda: (sums.zero_(), ops.bn_bwd_reduce(g, None, z, mean, invstd, sums, relu_scale=scale, relu_shift=shift)))
    t_bapp = timeit(lambda: ops.bn_bwd_apply(g, None, z, mean, invstd, gamma, sums, dz, None, dg, db, relu_scale=scale, relu_shift=shift))
    t_fused = timeit(lambda: ops.bn_bwd_fused(g, None, z, mean, invstd, gamma, sums, dz, barrier, None, dg, db, relu_scale=scale, relu_shift=shift))
    tr = [train(lambda: ops.bn_apply(z, scale, shift, y, relu=True)),
          train(lambda: (ops.bn_bwd_reduce(g, None, z, mean, invstd, sums, relu_scale=scale, relu_shift=shift),
                         ops.bn_bwd_apply(g, None, z, mean, invstd, gamma, sums, dz, None, dg, db, relu_scale=scale, relu_shift=shift))),
          train(lambda: ops.bn_bwd_fused(g, None, z, mean, invstd, gamma, sums, dz, barrier, None, dg, db, relu_scale=scale, relu_shift=shift))]
    print(f"[{n},{h},{w},{c}] bwd_fused {t_fused:6.1f} us vs two-pass {t_red + t_bapp:6.1f} us (cold, single launch) | back-to-back: apply {tr[0]:6.1f}  "
          f"two-pass bwd {tr[1]:6.1f}  fused bwd {tr[2]:6.1f} us")
    print(f"[{n},{h},{w},{c}] apply {t_apply:6.1f} us ({2*nbytes/t_apply/1e3:6.0f} GB/s)  bwd_reduce {t_red:6.1f} us ({2*nbytes/t_red/1e3:6.0f} GB/s)  "
          f"bwd_apply {t_bapp:6.1f} us ({3*nbytes/t_bapp/1e3:6.0f} GB/s)")
