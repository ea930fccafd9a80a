"""Teacher-forced per-layer parity INSIDE the assembled U-Net (the hard gate).

After one forward+backward of the B200 module, every kernel invocation is re-derived in fp32 PyTorch from the
tensors the engine itself stored for that layer (bf16-representable on both sides), so only accumulation order
and the single output rounding differ.  This localises wiring / indexing bugs that free-running end-to-end
comparisons blur (the random-init network amplifies rounding noise ~2x per stage)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def close(out, ref, what, tol=1e-2):
    out, ref = out.float(), ref.float()
    t = tol * ref.abs().max().item() + 1e-7
    err = (out - ref).abs()
    bad = err > (t + ref.abs() / 128)
    frac = float(bad.float().mean())
    assert frac <= 1e-4, f"{what}: max err {err.max().item():.4g} tol {t:.4g} bad {frac:.2e} max|ref| {ref.abs().max().item():.4g}"


def bn_train(z, gamma, beta):
    return F.batch_norm(z, None, None, gamma, beta, training=True, eps=1e-5)


@pytest.fixture(scope="module")
def run():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from hallucidet_b200.unet import Unet
    torch.manual_seed(123)
    m = Unet("resnet34", encoder_weights=None, in_channels=3, classes=3)
    m.segmentation_head[-1] = torch.nn.Sigmoid()
    m = m.cuda().train()
    x = torch.rand(2, 1, 64, 96, generator=torch.Generator().manual_seed(1)).repeat(1, 3, 1, 1).cuda()
    hal = m(x)
    dhal = torch.randn(hal.shape, generator=torch.Generator().manual_seed(2)).cuda()
    (hal * dhal).sum().backward()
    torch.cuda.synchronize()
    eng = next(iter(m._engines.values()))
    return m, eng, x, hal.detach(), dhal


def w16(conv):
    return conv.weight.detach().to(torch.bfloat16).float()


def test_forward_layers(run):
    m, eng, x, hal, _ = run
    st = eng.stem
    xq = x.to(torch.bfloat16).float()
    z = F.conv2d(xq, w16(st.conv), stride=2, padding=3)
    close(nchw(st.z), z, "stem conv")
    close(nchw(eng.a_stem), F.relu(bn_train(nchw(st.z), st.bn.weight, st.bn.bias)), "stem bn+relu")
    assert torch.equal(nchw(eng.p0), F.max_pool2d(nchw(eng.a_stem), 3, 2, 1))
    for blk in eng.blocks:
        c1, c2, cd = blk["c1"], blk["c2"], blk["cd"]
        xin = nchw(blk["x_in"])
        close(nchw(c1.z), F.conv2d(xin, w16(c1.conv), stride=c1.stride, padding=1), c1.name)
        close(nchw(blk["a1"]), F.relu(bn_train(nchw(c1.z), c1.bn.weight, c1.bn.bias)), c1.name + " bn")
        close(nchw(c2.z), F.conv2d(nchw(blk["a1"]), w16(c2.conv), padding=1), c2.name)
        if cd is not None:
            close(nchw(cd.z), F.conv2d(xin, w16(cd.conv), stride=cd.stride), cd.name)
            idn = bn_train(nchw(cd.z), cd.bn.weight, cd.bn.bias)
        else:
            idn = xin
        close(nchw(blk["out"]), F.relu(bn_train(nchw(c2.z), c2.bn.weight, c2.bn.bias) + idn), c2.name + " bn+res")
    xd = eng.blocks[-1]["out"]
    for i, d in enumerate(eng.dblocks):
        assert torch.equal(nchw(d["up"]), F.interpolate(nchw(xd), scale_factor=2, mode="nearest"))
        cat = nchw(d["up"]) if d["skip"] is None else torch.cat([nchw(d["up"]), nchw(d["skip"])], 1)
        c1, c2 = d["c1"], d["c2"]
        close(nchw(c1.z), F.conv2d(cat, w16(c1.conv), padding=1), c1.name)
        close(nchw(d["a1"]), F.relu(bn_train(nchw(c1.z), c1.bn.weight, c1.bn.bias)), c1.name + " bn")
        close(nchw(c2.z), F.conv2d(nchw(d["a1"]), w16(c2.conv), padding=1), c2.name)
        close(nchw(d["a2"]), F.relu(bn_train(nchw(c2.z), c2.bn.weight, c2.bn.bias)), c2.name + " bn")
        xd = d["a2"]
    head = m.segmentation_head[0]
    ref = torch.sigmoid(F.conv2d(nchw(xd), w16(head), head.bias.detach(), padding=1))
    assert (hal - ref).abs().max().item() < 2e-3


def bn_bwd_ref(z, gamma, beta, g, y_relu):
    """dz, dgamma, dbeta of train-mode BN at z for upstream gradient g masked by (y_relu > 0)."""
    zr = z.clone().requires_grad_(True)
    gm = gamma.detach().clone().requires_grad_(True)
    bt = beta.detach().clone().requires_grad_(True)
    y = bn_train(zr, gm, bt)
    gmask = g * (y_relu > 0)
    (y * gmask).sum().backward()
    return zr.grad, gm.grad, bt.grad, gmask


def test_backward_layers(run):
    m, eng, x, hal, dhal = run
    gb = eng.grad_bufs
    grads = {k: p.grad for k, p in m.named_parameters()}
    head = m.segmentation_head[0]
    dlog = dhal * hal * (1 - hal)
    close(nchw(eng.dlogits)[:, :3], dlog, "dlogits")
    assert torch.allclose(grads["segmentation_head.0.bias"], dlog.sum((0, 2, 3)), rtol=1e-3, atol=1e-3)
    dlq = nchw(eng.dlogits)[:, :3]
    a_last = eng.dblocks[-1]["a2"]
    close(grads["segmentation_head.0.weight"], torch.nn.grad.conv2d_weight(nchw(a_last), head.weight.shape, dlq, padding=1), "head wgrad", tol=2e-3)
    g = gb[("g", "head_in")]
    close(nchw(g), torch.nn.grad.conv2d_input(nchw(a_last).shape, w16(head), dlq, padding=1), "head dgrad")

    def check_conv_bn(l, g_in, y_relu, x_in_nchw, what, g_x_extra=None, g_x_buf=None):
        """BN backward + wgrad + (optionally) dgrad of layer l, all from the engine's own stored tensors."""
        dz_ref, dgam, dbet, _ = bn_bwd_ref(nchw(l.z), l.bn.weight, l.bn.bias, nchw(g_in), nchw(y_relu))
        dz = gb[("dz", l.name)]
        close(nchw(dz.view_as(l.z)), dz_ref, what + " bn bwd dz")
        assert torch.allclose(grads[l.bn_name + ".weight"], dgam, rtol=2e-2, atol=1e-2 * dgam.abs().max().item() + 1e-6), what + " dgamma"
        assert torch.allclose(grads[l.bn_name + ".bias"], dbet, rtol=2e-2, atol=1e-2 * dbet.abs().max().item() + 1e-6), what + " dbeta"
        dzq = nchw(dz.view_as(l.z))
        wref = torch.nn.grad.conv2d_weight(x_in_nchw, l.conv.weight.shape, dzq, stride=l.stride, padding=l.k // 2)
        close(grads[l.name + ".weight"], wref, what + " wgrad", tol=2e-3)
        return dzq

    # decoder, last block first
    for i in reversed(range(len(eng.dblocks))):
        d = eng.dblocks[i]
        c1, c2 = d["c1"], d["c2"]
        dz2 = check_conv_bn(c2, g, d["a2"], nchw(d["a1"]), c2.name)
        g_a1 = gb[("g", c2.name)]
        close(nchw(g_a1), torch.nn.grad.conv2d_input(nchw(d["a1"]).shape, w16(c2.conv), dz2, padding=1), c2.name + " dgrad")
        cat = nchw(d["up"]) if d["skip"] is None else torch.cat([nchw(d["up"]), nchw(d["skip"])], 1)
        dz1 = check_conv_bn(c1, g_a1, d["a1"], cat, c1.name)
        gcat = torch.nn.grad.conv2d_input(cat.shape, w16(c1.conv), dz1, padding=1)
        g_up = gb[("gup", i)]
        close(nchw(g_up), gcat[:, :d["cin"]], c1.name + " dgrad (up part)")
        if d["skip"] is not None:
            close(nchw(gb[("gskip", i)]), gcat[:, d["cin"]:], c1.name + " dgrad (skip part)")
        g = gb[("gdown", i)]
        close(nchw(g), F.avg_pool2d(nchw(g_up), 2) * 4, f"upsample bwd {i}")
    skip_of = {4: gb[("gskip", 0)], 3: gb[("gskip", 1)], 2: gb[("gskip", 2)]}
    for blk in reversed(eng.blocks):
        c1, c2, cd = blk["c1"], blk["c2"], blk["cd"]
        xin = nchw(blk["x_in"])
        dz2 = check_conv_bn(c2, g, blk["out"], nchw(blk["a1"]), c2.name)
        gmask = nchw(g) * (nchw(blk["out"]) > 0)
        g_a1 = gb[("g", c2.name)]
        close(nchw(g_a1), torch.nn.grad.conv2d_input(nchw(blk["a1"]).shape, w16(c2.conv), dz2, padding=1), c2.name + " dgrad")
        dz1 = check_conv_bn(c1, g_a1, blk["a1"], xin, c1.name)
        gx = torch.nn.grad.conv2d_input(xin.shape, w16(c1.conv), dz1, stride=c1.stride, padding=1)
        if cd is None:
            close(nchw(gb[("gm", c2.name)]), gmask, c2.name + " masked g")
            gx = gx + nchw(gb[("gm", c2.name)])
        else:
            dzd = check_conv_bn(cd, g, blk["out"], xin, cd.name)
            gx = gx + torch.nn.grad.conv2d_input(xin.shape, w16(cd.conv), dzd, stride=cd.stride) + nchw(skip_of[blk["li"]])
        g = gb[("gx", c1.name)]
        close(nchw(g), gx, c1.name + " input gradient", tol=2e-2)
    # stem
    st = eng.stem
    a = nchw(eng.a_stem).requires_grad_(True)
    (F.max_pool2d(a, 3, 2, 1) * nchw(g)).sum().backward()
    g_stem = gb[("g", "stem")]
    close(nchw(g_stem), a.grad + nchw(gb[("gskip", 3)]), "maxpool bwd + skip", tol=2e-2)
    xq = x.to(torch.bfloat16).float()
    dz_ref, dgam, dbet, _ = bn_bwd_ref(nchw(st.z), st.bn.weight, st.bn.bias, nchw(g_stem), nchw(eng.a_stem))
    dzs = gb[("dz", st.name)].view_as(st.z)
    close(nchw(dzs), dz_ref, "stem bn bwd")
    wref = torch.nn.grad.conv2d_weight(xq, st.conv.weight.shape, nchw(dzs), stride=2, padding=3)
    close(grads["encoder.conv1.weight"], wref, "stem wgrad", tol=2e-3)
