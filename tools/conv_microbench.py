"""Micro-benchmark of representative conv launches of config 2 (B=8): CUDA-event time per launch with an L2 flush
between repetitions, TFLOP/s and algorithmic GB/s.  Usage: python tools/conv_microbench.py [filter] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B = 8


def bf(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(torch.bfloat16)


CASES = [
    # name, kind, h, w, cin, cout, k, stride, flags
    ("fwd1x1_64_256_160", "fwd", 160, 160, 64, 256, 1, 1, "bias,relu"),
    ("fwd1x1_64_256_160_add", "fwd", 160, 160, 64, 256, 1, 1, "bias,relu,add"),
    ("fwd1x1_256_64_160", "fwd", 160, 160, 256, 64, 1, 1, "bias,relu"),
    ("fwd3x3_256_256_160", "fwd", 160, 160, 256, 256, 3, 1, "bias"),
    ("fwd3x3_64_64_128x160_stats", "fwd", 128, 160, 64, 64, 3, 1, "stats"),
    ("fwd3x3_128_128_64x80_stats", "fwd", 64, 80, 128, 128, 3, 1, "stats"),
    ("fwd3x3_256_256_32x40_stats", "fwd", 32, 40, 256, 256, 3, 1, "stats"),
    ("fwd3x3_512_512_16x20_stats", "fwd", 16, 20, 512, 512, 3, 1, "stats"),
    ("fwd3x3_64_64_128x160_statsfin", "fwd", 128, 160, 64, 64, 3, 1, "stats,fin"),
    ("fwd3x3_128_128_64x80_statsfin", "fwd", 64, 80, 128, 128, 3, 1, "stats,fin"),
    ("fwd3x3_256_256_32x40_statsfin", "fwd", 32, 40, 256, 256, 3, 1, "stats,fin"),
    ("fwd3x3_512_512_16x20_statsfin", "fwd", 16, 20, 512, 512, 3, 1, "stats,fin"),
    ("fwd3x3_16_16_512x640_stats", "fwd", 512, 640, 16, 16, 3, 1, "stats"),
    ("fwd3x3_32_16_512x640_stats", "fwd", 512, 640, 32, 16, 3, 1, "stats"),
    ("fwd3x3_32_32_256x320_stats", "fwd", 256, 320, 32, 32, 3, 1, "stats"),
    ("fwd3x3_16_16_512x640_head", "fwd", 512, 640, 16, 16, 3, 1, "bias"),
    ("fwd3x3_128_128_80_relu", "fwd", 80, 80, 128, 128, 3, 1, "bias,relu"),
    ("dgrad1x1_64to256_160_addmask", "dgrad", 160, 160, 256, 64, 1, 1, "add,mask"),
    ("dgrad3x3_64_64_160_mask", "dgrad", 160, 160, 64, 64, 3, 1, "mask"),
    ("dgrad3x3_64_64_128x160_mask", "dgrad", 128, 160, 64, 64, 3, 1, "mask"),
    ("dgrad3x3_64_64_128x160_addmask", "dgrad", 128, 160, 64, 64, 3, 1, "add,mask"),
    ("fwd3x3_64_64_160_relu", "fwd", 160, 160, 64, 64, 3, 1, "bias,relu"),
    ("fwd3x3_128_32_256x320_stats", "fwd", 256, 320, 128, 32, 3, 1, "stats"),
    ("dgrad3x3_256_256_40_mask", "dgrad", 40, 40, 256, 256, 3, 1, "mask"),
    ("dgrad3x3_16_16_512x640", "dgrad", 512, 640, 16, 16, 3, 1, ""),
    ("dgrad3x3_32_16_512x640", "dgrad", 512, 640, 32, 16, 3, 1, ""),
    ("dgrad3x3_32_32_256x320", "dgrad", 256, 320, 32, 32, 3, 1, ""),
    ("wgrad3x3_64_64_128x160", "wgrad", 128, 160, 64, 64, 3, 1, ""),
    ("wgrad3x3_128_128_64x80", "wgrad", 64, 80, 128, 128, 3, 1, ""),
    ("wgrad3x3_256_256_32x40", "wgrad", 32, 40, 256, 256, 3, 1, ""),
    ("wgrad3x3_512_512_16x20", "wgrad", 16, 20, 512, 512, 3, 1, ""),
    ("wgrad3x3_16_16_512x640", "wgrad", 512, 640, 16, 16, 3, 1, ""),
    ("wgrad3x3_32_16_512x640", "wgrad", 512, 640, 32, 16, 3, 1, ""),
    ("wgrad3x3_32_32_256x320", "wgrad", 256, 320, 32, 32, 3, 1, ""),
]


def main():
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    results = []
    for name, kind, h, w, cin, cout, k, s, flags in CASES:
        if filt and filt not in name:
            continue
        flags = set(f for f in flags.split(",") if f)
        wt = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).contiguous()
        pk = ops.PackedConv(cout, cin, k, dev).pack(wt)
        ho, wo = h // s, w // s
        x = bf(B, h, w, cin)
        y = bf(B, ho, wo, cout)
        keep = []                                          # conv_args holds raw pointers: keep every operand alive

        def own(t):
            keep.append(t)
            return t

        if kind == "fwd":
            fin = None
            if "fin" in flags:
                t = [own(torch.ones(cout, device=dev)) for _ in range(8)]
                fin = ops.bn_fin(B * ho * wo, t[0], t[1], 1e-5, 0.1, t[2], t[3], t[4], t[5], t[6], t[7],
                                 own(torch.zeros(1, dtype=torch.int32, device=dev)))
            args = ops.conv_args(x, y, pk.w_fwd, k=k, stride=s, bn_fin=fin,
                                 bias=own(torch.randn(cout, device=dev)) if "bias" in flags else None,
                                 add=own(bf(B, ho, wo, cout)) if "add" in flags else None, relu="relu" in flags,
                                 stats=own(torch.zeros(ops.conv_fwd_tiles(x, k, s), 2, cout, device=dev)) if "stats" in flags else None)
            fn = lambda: ops.conv_fwd(args)
            bytes_ = 2 * (x.numel() + y.numel() * (2 if "add" in flags else 1)) + 2 * wt.numel()
        elif kind == "dgrad":
            dx = bf(B, h, w, cin)
            args = ops.conv_args(y, dx, pk.w_dgrad, k=k, stride=s, add=own(bf(B, h, w, cin)) if "add" in flags else None,
                                 mask=own(bf(B, h, w, cin)) if "mask" in flags else None)
            fn = lambda: ops.conv_dgrad(args)
            bytes_ = 2 * (y.numel() + dx.numel() * (1 + ("add" in flags) + ("mask" in flags))) + 2 * wt.numel()
        else:
            dw = torch.zeros(cout, k * k, cin, device=dev)
            args = ops.conv_args(x, y, k=k, stride=s, dw=dw)
            fn = lambda: ops.conv_wgrad(args)
            bytes_ = 2 * (x.numel() + y.numel()) + 4 * wt.numel()
        flops = 2.0 * B * ho * wo * cin * cout * k * k
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        # GPU time without host launch cost: 10 back-to-back launches in a CUDA graph (operands L2-warm, PDL-chained)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                fn()
        g.replay()
        torch.cuda.synchronize()
        gts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            gts.append(e0.elapsed_time(e1) * 1e2)
        gts.sort()
        gus = gts[len(gts) // 2]
        results.append({"name": name, "us": us, "graph_us": gus, "tflops": flops / us / 1e6, "graph_tflops": flops / gus / 1e6,
                        "gbs": bytes_ / us / 1e3, "min_us": ts[0]})
        print(f"{name:34s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s  {bytes_ / us / 1e3:8.1f} GB/s (algorithmic) | in-graph {gus:7.1f} us {flops / gus / 1e6:8.1f} TFLOP/s")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/conv_microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
