"""In-kernel timeline of one conv launch (hd_conv_debug_timestamps): per-CTA %globaltimer stamps, relative to the first CTA's
start -> where a launch's time goes (launch ramp, dependency wait, operand latency, main loop, epilogue, tail).
Usage: python tools/conv_timeline.py [microbench-case-filter]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
NAMES = ["start", "deps ok", "last TMA", "1st operands", "last MMA", "last acc ready", "epilogue done", "exit"]
CASES = [
    # name, n, h, w, cin, cout, k, flags
    ("tiny 1 tile K=2304", 1, 8, 16, 256, 128, 3, ""),
    ("tiny 1 tile K=256 1x1", 1, 8, 16, 256, 128, 1, ""),
    ("fwd3x3_256_256_32x40 stats", 8, 32, 40, 256, 256, 3, "stats"),
    ("fwd3x3_256_256_32x40 plain", 8, 32, 40, 256, 256, 3, ""),
    ("fwd3x3_512_512_16x20 stats", 8, 16, 20, 512, 512, 3, "stats"),
    ("fwd3x3_128_128_64x80 stats", 8, 64, 80, 128, 128, 3, "stats"),
    ("fwd3x3_64_64_128x160 stats", 8, 128, 160, 64, 64, 3, "stats"),
    ("fwd3x3_64_64_160 bias relu", 8, 160, 160, 64, 64, 3, "bias"),
    ("fwd3x3_128_32_256x320 stats", 8, 256, 320, 128, 32, 3, "stats"),
    ("fwd1x1_64_256_160 bias relu", 8, 160, 160, 64, 256, 1, "bias"),
    ("fwd1x1_256_64_160 bias relu", 8, 160, 160, 256, 64, 1, "bias"),
    ("fwd3x3_256_256_160 bias", 8, 160, 160, 256, 256, 3, "bias"),
]
filt = sys.argv[1] if len(sys.argv) > 1 else ""
lib = _lib.load()
buf = torch.zeros(160, 8, dtype=torch.int64, device=dev)
for name, n, h, w, cin, cout, k, flags in CASES:
    if filt and filt not in name:
        continue
    x = (torch.randn(n, h, w, cin, device=dev) * 0.5).to(torch.bfloat16)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
    wt = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    pk = ops.PackedConv(cout, cin, k, dev).pack(wt)
    bias = torch.randn(cout, device=dev) if "bias" in flags else None
    stats = torch.zeros(ops.conv_fwd_tiles(x, k, 1, cout=cout), 2, cout, device=dev) if "stats" in flags else None
    args = ops.conv_args(x, y, pk.w_fwd, k=k, bias=bias, relu=bias is not None, stats=stats)
    for _ in range(3):
        ops.conv_fwd(args)
    torch.cuda.synchronize()
    for sk in (1, 0):
        os.environ["HD_STREAMK_RUNTIME"] = str(sk)
        ops.STREAMK = bool(sk)
        args = ops.conv_args(x, y, pk.w_fwd, k=k, bias=bias, relu=bias is not None, stats=stats)
        res = []
        for rep in range(5):
            buf.zero_()
            lib.hd_conv_debug_timestamps(ctypes.c_void_p(buf.data_ptr()))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_fwd(args)
            e1.record()
            torch.cuda.synchronize()
            lib.hd_conv_debug_timestamps(None)
            t = buf.cpu()
            live = t[:, 0] > 0
            t = t[live].double()
            t0 = t[:, 0].min()
            res.append((e0.elapsed_time(e1) * 1e3, (t - t0) / 1e3, int(live.sum())))
        res.sort(key=lambda r: r[0])
        us, rel, ctas = res[len(res) // 2]
        print(f"{name:32s} workspace(stream-K allowed)={sk} CTAs={ctas} event {us:6.1f} us; per-CTA stamps (us after first CTA start): min / median / max")
        for i, nm in enumerate(NAMES):
            col = rel[:, i][rel[:, i] >= 0]
            if len(col):
                print(f"    {nm:16s} {col.min():7.2f} {col.median():7.2f} {col.max():7.2f}")
