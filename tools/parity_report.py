"""Free-running parity report (GPU): B200 U-Net vs the fp32 oracle and the bf16-storage oracle, next to the oracle's own
bf16-storage noise floor (bf16-storage oracle vs fp32 oracle).  Writes gpurun_out/parity_report.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet as ou  # noqa: E402
from hallucidet_b200.unet import Unet  # noqa: E402


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def run(B, H, W):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(123)
    m = Unet("resnet34", encoder_weights=None, in_channels=3, classes=3)
    m.segmentation_head[-1] = torch.nn.Sigmoid()
    m = m.cuda().train()
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(1)).repeat(1, 3, 1, 1).cuda()
    gw = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(2)).cuda()
    hal = m(x)
    (hal * gw).sum().backward()
    g_mine = torch.cat([p.grad.flatten() for p in m.parameters()]).double()
    res = {}
    outs = {}
    for tag, q in (("fp32", ou._id), ("bf16_storage", ou.round_bf16)):
        st = {k: v.clone() for k, v in state.items()}
        params = [v.requires_grad_(True) for k, v in st.items() if ou.is_param(k)]
        h = ou.unet_forward(st, x, training=True, q=q)
        (h * gw).sum().backward()
        g = torch.cat([p.grad.flatten() for p in params]).double()
        outs[tag] = (h.detach(), g)
        e = (hal.detach() - h.detach()).abs()
        res[f"mine_vs_{tag}"] = {"hal_max": float(e.max()), "hal_mean": float(e.mean()), "grad_cos": cos(g_mine, g)}
    e = (outs["fp32"][0] - outs["bf16_storage"][0]).abs()
    res["noise_floor_bf16_storage_vs_fp32"] = {"hal_max": float(e.max()), "hal_mean": float(e.mean()),
                                               "grad_cos": cos(outs["fp32"][1], outs["bf16_storage"][1])}
    return res


if __name__ == "__main__":
    report = {}
    for shape in ((2, 64, 96), (2, 256, 320), (8, 512, 640)):
        report["x".join(map(str, shape))] = run(*shape)
        print(shape, json.dumps(report["x".join(map(str, shape))]))
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open("gpurun_out/parity_report.json", "w"), indent=1)
