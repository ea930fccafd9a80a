"""Stock-PyTorch GPU baselines of the config-2 train step on the same B200 (SURVEY.md 8d "GPU baseline to beat"):
the fp32 oracle restatement of the reference's modules through ATen / cuDNN, (i) fp32 with TF32 off, (ii) TF32 on +
cudnn.benchmark (what the reference scripts set, train_hallucidet.py:28-29), (iii) torch.autocast(bf16).
Forward + backward (+ Adam step) per iteration, batch 8, 512x640, S=640.  Tool only: not part of the product path."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet as ou, detector as odet, step as ostep

dev = torch.device("cuda", 0)
B = 8
ir, rgb, targets = ostep.synthetic_batch(B, 512, 640, seed=123, device=dev)
state = {k: v.to(dev) for k, v in ou.init_unet_state(123).items()}
det = odet.build_detector("fasterrcnn", seed=123).to(dev)
params = [v for k, v in state.items() if ou.is_param(k)]
opt = torch.optim.Adam(params, lr=1e-4, fused=True)


def step(autocast):
    opt.zero_grad(set_to_none=True)
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = ostep.train_step(state, det, ir, rgb, targets, size=640, detector_name="fasterrcnn", det_seed=7)
    else:
        out = ostep.train_step(state, det, ir, rgb, targets, size=640, detector_name="fasterrcnn", det_seed=7)
    torch.nn.utils.clip_grad_value_(params, 0.5)
    opt.step()
    return float(out["loss"])


res = {}
for name, tf32, bench, ac in (("fp32 (TF32 off)", False, False, False), ("fp32 + TF32 + cudnn.benchmark", True, True, False),
                              ("autocast bf16 + TF32 + cudnn.benchmark", True, True, True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = bench
    try:
        for _ in range(3):
            loss = step(ac)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 8
        for _ in range(n):
            loss = step(ac)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        res[name] = {"ms_per_step": ms, "images_per_s": B / ms * 1e3, "loss": loss}
        print(f"{name:42s} {ms:8.1f} ms/step  {B / ms * 1e3:7.1f} img/s  loss {loss:.4f}")
    except Exception as e:                      # e.g. an op without a bf16 kernel under autocast
        res[name] = {"error": repr(e)[:200]}
        print(name, "FAILED", repr(e)[:200])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gpu_baseline.json", "w"), indent=1)
