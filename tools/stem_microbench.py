"""GPU time of the stem kernels at the config-2 sizes (CUDA-graph replay of 10 launches, us per launch)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200 import ops
dev = torch.device("cuda", 0)

def graph_us(fn, n=10):
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

wt = torch.randn(64, 3, 7, 7, device=dev) / 12
for name, n, cin, h, w, dt in (("unet 1ch u8", 8, 1, 512, 640, torch.uint8), ("unet 1ch f32", 8, 1, 512, 640, torch.float32),
                               ("unet 3ch f32", 8, 3, 512, 640, torch.float32), ("backbone 3ch f32", 8, 3, 640, 640, torch.float32)):
    x = (torch.rand(n, cin, h, w, device=dev) * (255 if dt == torch.uint8 else 1)).to(dt)
    y = torch.empty(n, h // 2, w // 2, 64, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(ops.stem_fwd_rows(x), 2, 64, device=dev)
    bias = torch.randn(64, device=dev)
    t_plain = graph_us(lambda: ops.stem_fwd(x, wt, y, x_scale=1 / 255 if dt == torch.uint8 else 1.0, bias=bias, relu=True))
    t_stats = graph_us(lambda: ops.stem_fwd(x, wt, y, x_scale=1 / 255 if dt == torch.uint8 else 1.0, stats=stats))
    kp = 64 if cin == 1 else 160
    patches = torch.empty(1, 1, n * (h // 2) * (w // 2), kp, dtype=torch.bfloat16, device=dev)
    t_i2c = graph_us(lambda: ops.stem_im2col_1ch(x, patches, k_pad=kp) if cin == 1 else ops.stem_im2col(x, patches))
    out_mb = y.numel() * 2 / 1e6
    print(f"{name:18s} fused fwd (bias+relu) {t_plain:6.1f} us  (+stats) {t_stats:6.1f} us  [{out_mb:.0f} MB out -> {out_mb / 6545 * 1e3:.1f} us at HBM peak]   im2col {t_i2c:6.1f} us")
