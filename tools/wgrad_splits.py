"""time the tcgen05 wgrad for one shape across split_k values."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200 import ops
dev = torch.device("cuda", 0)
B = 8
def bf(*s): return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (h, w, cin, cout) in [(128, 160, 64, 64), (64, 80, 128, 128), (32, 40, 256, 256)]:
    x, y = bf(B, h, w, cin), bf(B, h, w, cout)
    dw = torch.zeros(cout, 9, cin, device=dev)
    for sk in (4, 8, 16, 33, 66, 132):
        args = ops.conv_args(x, y, k=3, stride=1, dw=dw, split_k=sk)
        for _ in range(3): ops.conv_wgrad(args)
        ts = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.conv_wgrad(args); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        tiles = B * h * w // 128
        print(f"{h}x{w} {cin}->{cout} split_k={sk:4d} ctas={sk*9:5d} tiles/cta={tiles/sk:6.1f} {ts[len(ts)//2]:8.1f} us")
