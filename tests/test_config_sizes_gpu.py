"""Parity AT THE BASELINE.json CONFIGURATION SIZES.

Kernel selection inside the library is shape dependent (128- vs 256-wide N tiles by wave count, split-K factors, stride-2
phase views, the narrow-layer kernels, per-CTA statistics rows), so small-shape tests do not enter the code paths the
benchmark runs.  Here:

  * one full train step of config 2 (B=8, 512x640 IR, S=640, Faster R-CNN) and of config 4 (RetinaNet) runs through the
    B200 path with every conv launch RECORDED (ops.RECORD); its loss is gated against the fp32 oracle run on the same GPU
    on the same weights / inputs / detector seed: |loss - oracle| <= 1 % (north star);
  * every DISTINCT recorded launch signature (operand shapes, filter, stride, fused epilogue operands, output kinds) is then
    replayed teacher-forced: random bf16-representable operands of exactly that shape through the C ABI, against plain fp32
    PyTorch (conv2d / conv2d_input / conv2d_weight + the epilogue arithmetic).  Tolerances as tests/test_kernels_gpu.py:
    bf16 outputs |err| <= 1e-2 max|ref| + |ref|/256, fp32 outputs / weight gradients rtol 2e-3 of max|ref|, BN statistics
    rtol 2e-3.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _bf(t):
    return t.to(torch.bfloat16)


def _rnd(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return _bf(torch.randn(*shape, generator=g, device="cuda") * scale)


def _nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def _bf16_close(out, ref):
    out, ref = out.float(), ref.float()
    tol = 1e-2 * ref.abs().max().item() + 1e-6
    err = (out - ref).abs()
    bad = err > (tol + ref.abs() / 256)
    return None if not bad.any() else f"bf16 out: max err {err.max().item():.4g} tol {tol:.4g}, {int(bad.sum())}/{bad.numel()} bad"


def _f32_close(out, ref, rtol=2e-3):
    err = (out.float() - ref.float()).abs().max().item()
    lim = rtol * ref.abs().max().item() + 1e-4
    return None if err <= lim else f"fp32 out: max err {err:.4g} > {lim:.4g}"


def replay_launch(sig, seed=0):
    """Run one recorded launch signature on random operands and compare with fp32 PyTorch.  Returns None or an error text."""
    from hallucidet_b200 import ops as o
    (kind, x0s, x1c, y0s, y1c, k, s, has_bias, has_add, has_mask, relu, sigmoid, has_stats, has_f32, f32c, f32_nhwc,
     store_bf16, phase_mask, inplace) = sig
    n, h, w, c0 = x0s
    pad = k // 2
    gcpu = torch.Generator().manual_seed(seed + 17)
    if kind == "fwd":
        cin, cout = c0 + x1c, y0s[3]
        ho, wo = y0s[1], y0s[2]
        x0 = _rnd((n, h, w, c0), seed + 1)
        x1 = _rnd((n, h, w, x1c), seed + 2) if x1c else None
        wt = _bf(torch.randn(cout, cin, k, k, generator=gcpu) / (cin * k * k) ** 0.5).float().cuda()
        pk = o.PackedConv(cout, cin, k, "cuda", need_dgrad=False).pack(wt)
        y = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
        bias = torch.randn(cout, device="cuda") * 0.1 if has_bias else None
        add = _rnd((n, ho, wo, cout), seed + 3) if has_add else None
        mask = _rnd((n, ho, wo, cout), seed + 4) if has_mask else None
        stats = None
        if has_stats:
            rows = o.conv_fwd_tiles(x0, k, s, cout=None if x1 is not None else cout)
            stats = torch.full((rows, 2, cout), float("nan"), device="cuda")
        f32 = None
        if has_f32:
            f32 = (torch.zeros(n, ho, wo, f32c, device="cuda").permute(0, 3, 1, 2) if f32_nhwc
                   else torch.zeros(n, f32c, ho, wo, device="cuda"))
        o.conv_fwd(o.conv_args(x0, y, pk.w_fwd, k=k, stride=s, x1=x1, bias=bias, add=add, mask=mask, relu=bool(relu),
                               sigmoid=bool(sigmoid), stats=stats, out_f32=f32, out_f32_channels=f32c, store_bf16=bool(store_bf16),
                               out_f32_nhwc=bool(f32_nhwc)))
        torch.cuda.synchronize()
        xin = _nchw(x0) if x1 is None else torch.cat([_nchw(x0), _nchw(x1)], 1)
        ref = F.conv2d(xin, wt, bias, stride=s, padding=pad)
        del xin
    elif kind == "dgrad":
        cout = c0                                           # x0 = dY
        cin = y0s[3] + y1c
        H, W = y0s[1], y0s[2]
        dy = _rnd((n, h, w, cout), seed + 1)
        wt = _bf(torch.randn(cout, cin, k, k, generator=gcpu) / (cout * k * k) ** 0.5).float().cuda()
        pk = o.PackedConv(cout, cin, k, "cuda").pack(wt)
        y = torch.full((n, H, W, y0s[3]), float("nan"), dtype=torch.bfloat16, device="cuda")
        y1 = torch.full((n, H, W, y1c), float("nan"), dtype=torch.bfloat16, device="cuda") if y1c else None
        bias = torch.randn(cin, device="cuda") * 0.1 if has_bias else None
        add = _rnd((n, H, W, cin), seed + 3) if has_add else None
        if inplace:
            y.copy_(add)
        mask = _rnd((n, H, W, cin), seed + 4) if has_mask else None
        o.conv_dgrad(o.conv_args(dy, y, pk.w_dgrad, k=k, stride=s, y1=y1, bias=bias, add=(y if inplace else add), mask=mask,
                                 relu=bool(relu), phase_mask=phase_mask))
        torch.cuda.synchronize()
        ref = torch.nn.grad.conv2d_input((n, cin, H, W), wt, _nchw(dy), stride=s, padding=pad)
        if bias is not None:
            ref = ref + bias.view(1, -1, 1, 1)
        stats, f32 = None, None
        if y1 is not None:
            y = torch.cat([y, y1], 3)
    else:                                                   # wgrad: x = conv input, y0 = dY
        cin, cout = c0 + x1c, y0s[3]
        x0 = _rnd((n, h, w, c0), seed + 1)
        x1 = _rnd((n, h, w, x1c), seed + 2) if x1c else None
        dy = _rnd((n, y0s[1], y0s[2], cout), seed + 4, scale=0.05)
        dw = torch.zeros(cout, k * k, cin, device="cuda")
        o.conv_wgrad(o.conv_args(x0, dy, k=k, stride=s, x1=x1, dw=dw))
        torch.cuda.synchronize()
        xin = _nchw(x0) if x1 is None else torch.cat([_nchw(x0), _nchw(x1)], 1)
        ref = torch.nn.grad.conv2d_weight(xin, (cout, cin, k, k), _nchw(dy), stride=s, padding=pad)
        return _f32_close(dw, ref.permute(0, 2, 3, 1).reshape(cout, k * k, cin))
    if has_add:
        ref = ref + _nchw(add)
    if relu:
        ref = F.relu(ref)
    if has_mask:
        ref = ref * (_nchw(mask) > 0)
    errs = []
    if store_bf16:
        e = _bf16_close(_nchw(y), ref)
        if e:
            errs.append(e)
    if f32 is not None:
        r = torch.sigmoid(ref) if sigmoid else ref
        e = _f32_close(f32[:, :f32c], r[:, :f32c], rtol=2e-3 if not sigmoid else 4e-3)
        if e:
            errs.append(e)
    if stats is not None:
        yq = _nchw(y)
        ssum = stats.double().sum(0)
        for i, want in enumerate((yq.double().sum((0, 2, 3)), (yq.double() ** 2).sum((0, 2, 3)))):
            if not torch.allclose(ssum[i], want, rtol=2e-3, atol=1e-2 * max(1.0, float(want.abs().max()) * 1e-3)):
                errs.append(f"stats[{i}] max err {(ssum[i] - want).abs().max().item():.4g}")
    return "; ".join(errs) if errs else None


def _one_step(detector_name, B=8, H=512, W=640, S=640):
    """One recorded train step through the B200 modules + the fp32 oracle's loss for the same weights / inputs."""
    from hallucidet_b200 import ops
    from hallucidet_b200.synthetic import synthetic_batch
    from hallucidet_b200.train import HalluciDetTrainer
    from oracle import detector as odet, step as ostep
    dev = torch.device("cuda", 0)
    ir, rgb, targets = synthetic_batch(B, H, W, seed=123, device=dev)
    det_cpu = odet.build_detector(detector_name, seed=123)
    odet.randomize_bn_stats(det_cpu, seed=7)
    weights = {"pixel_rgb": 1.0, "pixel_ir": 0.5}
    tr = HalluciDetTrainer(detector_name=detector_name, size=S, pixel="mse", weights=weights, seed=123, device=dev,
                           detector_state=det_cpu.state_dict())
    state = {k: v.detach().clone() for k, v in tr.encoder_decoder.state_dict().items()}
    tr.encoder_decoder.train()
    ops.RECORD = []
    try:
        out = tr.forward_step(rgb, targets, ir, targets, det_seed=7)
        out["total"].backward()
        torch.cuda.synchronize()
    finally:
        rec, ops.RECORD = ops.RECORD, None
    mine = float(out["total"].detach())
    grads = {k: p.grad.detach().clone() for k, p in tr.encoder_decoder.named_parameters()}
    hal = out["hal"].detach().clone()
    del tr, out
    torch.cuda.empty_cache()
    ref = ostep.train_step(state, det_cpu.to(dev), ir, rgb, targets, size=S, detector_name=detector_name, pixel="mse",
                           weights=weights, det_seed=7)
    want = float(ref["loss"])
    keys = [k for k in grads if k in ref["grads"]]
    a = torch.cat([grads[k].double().flatten() for k in keys])
    b = torch.cat([ref["grads"][k].double().flatten() for k in keys])
    cosine = float((a @ b) / (a.norm() * b.norm()))
    herr = (hal - ref["hal"]).abs()
    info = {"loss": mine, "oracle_loss": want, "rel": abs(mine - want) / abs(want), "hal_max": float(herr.max()),
            "hal_mean": float(herr.mean()), "grad_cos_vs_fp32": cosine, "launches": len(rec), "distinct": len(set(rec))}
    del ref
    torch.cuda.empty_cache()
    return info, sorted(set(rec), key=repr)


@pytest.mark.parametrize("detector_name", ["fasterrcnn", "retinanet"])
def test_config_size_step_loss_and_launch_set(detector_name):
    """BASELINE.json config 2 (Faster R-CNN) / config 4 (RetinaNet) per-GPU step: B=8, 512x640 -> S=640."""
    info, launch_set = _one_step(detector_name)
    print(f"\n[config size, {detector_name}] " + " ".join(f"{k}={v:.5g}" if isinstance(v, float) else f"{k}={v}" for k, v in info.items()))
    assert info["rel"] <= 1e-2, f"train-step loss differs from the fp32 oracle by {info['rel']:.4%}"
    assert info["hal_mean"] <= 3e-2
    failures = []
    for i, sig in enumerate(launch_set):
        err = replay_launch(sig, seed=i)
        if err:
            failures.append(f"{sig}: {err}")
        torch.cuda.empty_cache()
    print(f"[config size, {detector_name}] replayed {len(launch_set)} distinct launch signatures, {len(failures)} failed")
    assert not failures, "\n".join(failures[:20])
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        import json
        json.dump({"step": info, "launch_signatures": [repr(s) for s in launch_set]},
                  open(os.path.join(out_dir, f"r2_config_size_parity_{detector_name}.json"), "w"), indent=1)
