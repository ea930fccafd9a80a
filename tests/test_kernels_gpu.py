"""Per-kernel parity on the GPU: every C-ABI entry point against a plain fp32 PyTorch restatement of the same
operator, fed bf16-representable inputs (so only accumulation order and the single output rounding differ).
Tolerances: bf16 outputs  |err| <= 1e-2 * max|ref| + 2^-8 |ref|;  fp32 outputs (wgrad, reductions) rtol 2e-3."""
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
    from hallucidet_b200 import _lib
    _lib.check(_lib.load().hd_device_ok(), "hd_device_ok")


def ops():
    from hallucidet_b200 import ops as o
    return o


def rnd(*shape, seed=0, scale=1.0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(device)


def nchw(x):            # NHWC bf16 -> NCHW fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def nhwc(x):            # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def assert_close_bf16(out, ref, what=""):
    out, ref = out.float(), ref.float()
    tol = 1e-2 * ref.abs().max().item() + 1e-6
    err = (out - ref).abs()
    bad = err > (tol + ref.abs() / 256)
    assert not bad.any(), f"{what}: max err {err.max().item():.4g} (tol {tol:.4g}), {int(bad.sum())}/{bad.numel()} bad, max ref {ref.abs().max().item():.4g}"


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, cin2
    (2, 16, 20, 64, 64, 3, 1, 0),
    (1, 32, 40, 128, 256, 3, 1, 0),
    (2, 8, 8, 256, 64, 1, 1, 0),
    (2, 16, 16, 16, 16, 3, 1, 0),
    (1, 16, 16, 32, 32, 3, 1, 0),
    (1, 64, 64, 32, 16, 3, 1, 0),
    (2, 20, 44, 16, 32, 3, 1, 0),       # narrow-layer kernel, partial 8x32 tiles
    (3, 9, 70, 32, 32, 3, 1, 0),
    (2, 64, 80, 64, 256, 3, 1, 0),      # 160 tiles of 128 channels > 148 SMs -> 256-wide N tiles
    (2, 64, 84, 64, 512, 3, 1, 0),
    (2, 16, 24, 64, 128, 3, 2, 0),
    (2, 16, 24, 64, 128, 1, 2, 0),
    (1, 16, 16, 64, 32, 3, 1, 64),
    (1, 8, 10, 128, 64, 3, 1, 64),
    (1, 20, 20, 512, 512, 3, 1, 0),
    (1, 10, 10, 2048, 256, 1, 1, 0),
    (1, 6, 6, 256, 256, 3, 2, 0),
    # stream-K (tiles under-fill the SMs, long K): config-2 deep-layer shapes and a ragged one
    (8, 32, 40, 256, 256, 3, 1, 0),
    (8, 16, 20, 512, 512, 3, 1, 0),
    (8, 64, 80, 128, 128, 3, 1, 0),
    (8, 32, 40, 512, 256, 3, 1, 256),
    (3, 20, 20, 512, 512, 3, 1, 0),
    (5, 40, 40, 1024, 256, 1, 1, 0),
    # halo mode (8x16 tiles, three 8x18 boxes per k-block, resident weights): ragged edges, many tiles per CTA, 2 k-blocks
    (2, 23, 37, 64, 64, 3, 1, 0),
    (8, 128, 160, 64, 64, 3, 1, 0),
    (3, 50, 30, 64, 32, 3, 1, 0),
    (2, 33, 20, 64, 16, 3, 1, 0),
    (4, 64, 72, 64, 32, 3, 1, 64),
    (1, 17, 9, 128, 32, 3, 1, 0),
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,cin2", CONV_CASES)
def test_conv_fwd(n, h, w, cin, cout, k, s, cin2):
    o = ops()
    x0 = rnd(n, h, w, cin, seed=1)
    x1 = rnd(n, h, w, cin2, seed=2) if cin2 else None
    ct = cin + cin2
    wt = (torch.randn(cout, ct, k, k, generator=torch.Generator().manual_seed(3)) / (ct * k * k) ** 0.5).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(cout, ct, k, "cuda").pack(wt)
    y = torch.full((n, h // s, w // s, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    o.conv_fwd(o.conv_args(x0, y, pk.w_fwd, k=k, stride=s, x1=x1))
    torch.cuda.synchronize()
    xin = nchw(x0) if x1 is None else torch.cat([nchw(x0), nchw(x1)], 1)
    ref = F.conv2d(xin, wt, stride=s, padding=k // 2)
    assert_close_bf16(nchw(y), ref, "conv_fwd")


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 24, 40, 64, 128), (2, 64, 80, 64, 256)])
def test_conv_fwd_epilogue_bias_add_relu_stats(n, h, w, cin, cout):
    o = ops()
    x = rnd(n, h, w, cin, seed=1)
    res = rnd(n, h, w, cout, seed=5)
    bias = torch.randn(cout, device="cuda")
    wt = (torch.randn(cout, cin, 3, 3, generator=torch.Generator().manual_seed(3)) / (cin * 9) ** 0.5).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(cout, cin, 3, "cuda").pack(wt)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    stats = torch.full((o.conv_fwd_tiles(x, 3, 1), 2, cout), float("nan"), device="cuda")   # rows are written, not accumulated
    f32 = torch.zeros(n, cout, h, w, device="cuda")
    o.conv_fwd(o.conv_args(x, y, pk.w_fwd, k=3, bias=bias, add=res, relu=True, stats=stats, out_f32=f32, out_f32_channels=cout))
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(nchw(x), wt, bias, padding=1) + nchw(res))
    assert_close_bf16(nchw(y), ref, "epilogue")
    assert torch.allclose(f32, ref, rtol=1e-3, atol=1e-3 * ref.abs().max().item())
    yq = nchw(y)
    s = stats.sum(0)
    assert torch.allclose(s[0], yq.sum((0, 2, 3)), rtol=2e-3, atol=1e-2)
    assert torch.allclose(s[1], (yq * yq).sum((0, 2, 3)), rtol=2e-3, atol=1e-2)


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 16), (16, 32), (32, 32)])
def test_conv_fwd_narrow_epilogue_bias_relu_stats(cin, cout):
    """The 16/32-channel 3x3 layers (decoder blocks 3-4, head) run on the halo-patch kernel: same epilogue contract."""
    o = ops()
    n, h, w = 2, 21, 75
    x = rnd(n, h, w, cin, seed=1)
    bias = torch.randn(cout, device="cuda")
    wt = (torch.randn(cout, cin, 3, 3, generator=torch.Generator().manual_seed(3)) / (cin * 9) ** 0.5).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(cout, cin, 3, "cuda").pack(wt)
    y = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    stats = torch.full((o.conv_fwd_tiles(x, 3, 1, cout=cout), 2, cout), float("nan"), device="cuda")
    o.conv_fwd(o.conv_args(x, y, pk.w_fwd, k=3, bias=bias, relu=True, stats=stats))
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(nchw(x), wt, bias, padding=1))
    assert_close_bf16(nchw(y), ref, "narrow epilogue")
    yq = nchw(y)
    s = stats.sum(0)
    assert torch.allclose(s[0], yq.sum((0, 2, 3)), rtol=2e-3, atol=1e-2)
    assert torch.allclose(s[1], (yq * yq).sum((0, 2, 3)), rtol=2e-3, atol=1e-2)
    # run-to-run deterministic statistics (fixed-order reductions, no atomics); a buffer sized without the channel hint
    # (more rows than CTAs) has its surplus rows zeroed
    stats2 = torch.full((o.conv_fwd_tiles(x, 3, 1), 2, cout), float("nan"), device="cuda")
    assert stats2.shape[0] >= stats.shape[0]
    o.conv_fwd(o.conv_args(x, y, pk.w_fwd, k=3, bias=bias, relu=True, stats=stats2))
    torch.cuda.synchronize()
    assert torch.equal(stats, stats2[:stats.shape[0]])
    assert float(stats2[stats.shape[0]:].abs().sum()) == 0.0
    # fp32 NCHW side output (the head's epilogue), without the bf16 store
    f32 = torch.zeros(n, cout, h, w, device="cuda")
    y2 = torch.zeros_like(y)
    o.conv_fwd(o.conv_args(x, y2, pk.w_fwd, k=3, bias=bias, relu=True, out_f32=f32, out_f32_channels=cout, store_bf16=False))
    torch.cuda.synchronize()
    assert torch.allclose(f32, ref, rtol=1e-3, atol=1e-3 * ref.abs().max().item())
    assert float(y2.float().abs().sum()) == 0.0


def test_conv_fwd_head_sigmoid_nchw():
    o = ops()
    n, h, w = 2, 32, 64
    x = rnd(n, h, w, 16, seed=1)
    wt = (torch.randn(3, 16, 3, 3, generator=torch.Generator().manual_seed(3)) / 12).to(torch.bfloat16).float().cuda()
    bias = torch.randn(3, device="cuda")
    pk = o.PackedConv(3, 16, 3, "cuda").pack(wt)
    out = torch.zeros(n, 3, h, w, device="cuda")
    ydummy = torch.empty(n, h, w, 16, dtype=torch.bfloat16, device="cuda")
    bias16 = torch.zeros(16, device="cuda")
    bias16[:3] = bias
    o.conv_fwd(o.conv_args(x, ydummy, pk.w_fwd, k=3, bias=bias16, sigmoid=True, out_f32=out, out_f32_channels=3, store_bf16=False))
    torch.cuda.synchronize()
    ref = torch.sigmoid(F.conv2d(nchw(x), wt, bias, padding=1))
    assert (out - ref).abs().max().item() < 2e-3


DGRAD_CASES = [
    # n, h, w (input dims), cin, cout, k, stride, split (cin of y1)
    (2, 16, 20, 64, 64, 3, 1, 0),
    (1, 32, 40, 256, 128, 3, 1, 0),
    (2, 8, 8, 64, 256, 1, 1, 0),
    (2, 16, 16, 16, 16, 3, 1, 0),
    (1, 64, 64, 32, 16, 3, 1, 0),
    (2, 20, 44, 32, 32, 3, 1, 0),       # narrow-layer kernel, partial 8x32 tiles
    (3, 9, 70, 16, 32, 3, 1, 0),
    (2, 64, 80, 256, 64, 3, 1, 0),      # 256-wide N tiles (dX has 256 channels)
    (2, 16, 24, 64, 128, 3, 2, 0),
    (1, 16, 16, 128, 64, 3, 1, 64),     # dX split into (64 | 64)
    (1, 16, 20, 192, 64, 3, 1, 64),     # (128 | 64)
    (8, 32, 40, 256, 256, 3, 1, 0),     # stream-K
    (8, 16, 20, 512, 512, 3, 1, 0),
    (8, 32, 40, 768, 256, 3, 1, 256),   # stream-K with a split gradient (512 | 256)
    # halo mode
    (2, 23, 37, 64, 64, 3, 1, 0),
    (8, 128, 160, 64, 64, 3, 1, 0),
    (3, 50, 30, 32, 64, 3, 1, 0),       # dX has 32 channels, dY 64
    (1, 17, 9, 32, 128, 3, 1, 0),       # two k-blocks
]


@pytest.mark.parametrize("n,h,w,c,use_add", [(2, 40, 48, 64, True), (2, 40, 48, 64, False), (8, 128, 160, 64, True), (3, 21, 19, 64, False)])
def test_conv_dgrad_halo_add_mask(n, h, w, c, use_add):
    """64 -> 64 3x3 input gradient with the fused ReLU mask (+ residual gradient): the halo-mode main loop under the
    register-store epilogue with its add / mask ring (one-deep next to the resident weights when both operands are fused)."""
    o = ops()
    dy = rnd(n, h, w, c, seed=2)
    wt = (torch.randn(c, c, 3, 3, generator=torch.Generator().manual_seed(4)) / (c * 9) ** 0.5).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(c, c, 3, "cuda").pack(wt)
    base = rnd(n, h, w, c, seed=7) if use_add else None
    msk = rnd(n, h, w, c, seed=9)
    out = torch.full((n, h, w, c), float("nan"), dtype=torch.bfloat16, device="cuda")
    o.conv_dgrad(o.conv_args(dy, out, pk.w_dgrad, k=3, add=base, mask=msk))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((n, c, h, w), wt, nchw(dy), padding=1)
    if use_add:
        ref = ref + nchw(base)
    ref = ref * (nchw(msk) > 0)
    assert_close_bf16(nchw(out), ref, "halo dgrad add/mask")


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,split", DGRAD_CASES)
def test_conv_dgrad(n, h, w, cin, cout, k, s, split):
    o = ops()
    dy = rnd(n, h // s, w // s, cout, seed=1)
    wt = (torch.randn(cout, cin, k, k, generator=torch.Generator().manual_seed(3)) / (cout * k * k) ** 0.5).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(cout, cin, k, "cuda").pack(wt)
    c0 = cin - split
    dx0 = torch.full((n, h, w, c0), float("nan"), dtype=torch.bfloat16, device="cuda")
    dx1 = torch.full((n, h, w, split), float("nan"), dtype=torch.bfloat16, device="cuda") if split else None
    o.conv_dgrad(o.conv_args(dy, dx0, pk.w_dgrad, k=k, stride=s, y1=dx1))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt, nchw(dy), stride=s, padding=k // 2)
    got = nchw(dx0) if dx1 is None else torch.cat([nchw(dx0), nchw(dx1)], 1)
    assert_close_bf16(got, ref, "conv_dgrad")


def test_conv_dgrad_1x1_s2_inplace_add_and_mask():
    o = ops()
    n, h, w, cin, cout = 2, 16, 24, 64, 128
    dy = rnd(n, h // 2, w // 2, cout, seed=1)
    wt = (torch.randn(cout, cin, 1, 1, generator=torch.Generator().manual_seed(3)) / cout ** 0.5).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(cout, cin, 1, "cuda").pack(wt)
    base = rnd(n, h, w, cin, seed=7)
    dx = base.clone()
    o.conv_dgrad(o.conv_args(dy, dx, pk.w_dgrad, k=1, stride=2, add=dx))
    torch.cuda.synchronize()
    ref = nchw(base) + torch.nn.grad.conv2d_input((n, cin, h, w), wt, nchw(dy), stride=2, padding=0)
    assert_close_bf16(nchw(dx), ref, "dgrad 1x1 s2 in-place add")
    # stride-1 with add + mask
    dy1 = rnd(n, h, w, cout, seed=2)
    wt3 = (torch.randn(cout, cin, 3, 3, generator=torch.Generator().manual_seed(4)) / (cout * 9) ** 0.5).to(torch.bfloat16).float().cuda()
    pk3 = o.PackedConv(cout, cin, 3, "cuda").pack(wt3)
    msk = rnd(n, h, w, cin, seed=9)
    out = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device="cuda")
    o.conv_dgrad(o.conv_args(dy1, out, pk3.w_dgrad, k=3, add=base, mask=msk))
    torch.cuda.synchronize()
    ref = (torch.nn.grad.conv2d_input((n, cin, h, w), wt3, nchw(dy1), padding=1) + nchw(base)) * (nchw(msk) > 0)
    assert_close_bf16(nchw(out), ref, "dgrad add+mask")


WGRAD_CASES = [
    # n, h, w (input dims), cin, cout, k, stride, cin2
    (2, 16, 20, 64, 64, 3, 1, 0),
    (2, 32, 40, 128, 256, 3, 1, 0),
    (2, 8, 8, 256, 64, 1, 1, 0),
    (2, 32, 32, 16, 16, 3, 1, 0),
    (1, 64, 64, 32, 16, 3, 1, 0),
    (2, 32, 32, 32, 32, 3, 1, 0),
    (2, 20, 44, 16, 32, 3, 1, 0),       # narrow-layer kernel, partial 8x32 tiles
    (3, 9, 70, 32, 32, 3, 1, 0),
    (2, 16, 24, 64, 128, 3, 2, 0),
    (2, 16, 24, 64, 128, 1, 2, 0),
    (1, 16, 16, 64, 32, 3, 1, 64),
    (1, 16, 20, 512, 256, 3, 1, 256),
    (2, 16, 20, 512, 512, 3, 1, 0),
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,cin2", WGRAD_CASES)
def test_conv_wgrad(n, h, w, cin, cout, k, s, cin2):
    o = ops()
    x0 = rnd(n, h, w, cin, seed=1)
    x1 = rnd(n, h, w, cin2, seed=2) if cin2 else None
    ct = cin + cin2
    dy = rnd(n, h // s, w // s, cout, seed=4)
    dw = torch.zeros(cout, k * k, ct, device="cuda")
    o.conv_wgrad(o.conv_args(x0, dy, k=k, stride=s, x1=x1, dw=dw))
    g = torch.empty(cout, ct, k, k, device="cuda")
    o.unpack_wgrad(dw, g, cout, ct, k, ct, k * k * ct)
    torch.cuda.synchronize()
    xin = nchw(x0) if x1 is None else torch.cat([nchw(x0), nchw(x1)], 1)
    ref = torch.nn.grad.conv2d_weight(xin, (cout, ct, k, k), nchw(dy), stride=s, padding=k // 2)
    err = (g - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-4, f"wgrad err {err} vs max {ref.abs().max().item()}"


def test_pack_weight_layouts():
    o = ops()
    cout, cin, k = 24, 40, 3
    wt = torch.randn(cout, cin, k, k, device="cuda").to(torch.bfloat16).float()
    sc = torch.rand(cout, device="cuda").to(torch.bfloat16).float() + 0.5
    pk = o.PackedConv(cout, cin, k, "cuda", need_t=True).pack(wt, sc)
    torch.cuda.synchronize()
    ws = (wt * sc[:, None, None, None]).to(torch.bfloat16)
    fwd = ws.permute(0, 2, 3, 1).reshape(cout, k * k * cin)
    assert torch.equal(pk.w_fwd[:cout], fwd) and (pk.w_fwd[cout:] == 0).all()
    dg = ws.permute(1, 2, 3, 0).reshape(cin, k * k * cout)
    assert torch.equal(pk.w_dgrad[:cin], dg) and (pk.w_dgrad[cin:] == 0).all()
    assert torch.equal(pk.w_t, pk.w_fwd.t())


@pytest.mark.parametrize("h,w", [(32, 64), (44, 100)])
def test_stem_im2col_gemm_col2im(h, w):
    o = ops()
    n = 2
    x = torch.rand(n, 3, h, w, device="cuda").to(torch.bfloat16).float()
    wt = (torch.randn(64, 3, 7, 7, generator=torch.Generator().manual_seed(3)) / 12).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(64, 3, 7, "cuda", need_dgrad=False, need_t=True, k_pad=o.STEM_KPAD).pack(wt)
    ho, wo = h // 2, w // 2
    patches = torch.empty(1, 1, n * ho * wo, o.STEM_KPAD, dtype=torch.bfloat16, device="cuda")
    patches.fill_(float("nan"))
    o.stem_im2col(x, patches)
    torch.cuda.synchronize()
    # exact patch matrix: k = (r*7 + s)*3 + c, zero padding of the image border and of the 147..159 tail
    unf = F.unfold(x, 7, padding=3, stride=2).view(n, 3, 49, ho * wo).permute(0, 3, 2, 1).reshape(n * ho * wo, 147)
    assert torch.equal(patches.view(-1, o.STEM_KPAD)[:, :147].float(), unf)
    assert float(patches.view(-1, o.STEM_KPAD)[:, 147:].float().abs().sum()) == 0.0
    y = torch.empty(1, 1, n * ho * wo, 64, dtype=torch.bfloat16, device="cuda")
    o.conv_fwd(o.conv_args(patches, y, pk.w_fwd, k=1))
    torch.cuda.synchronize()
    ref = F.conv2d(x, wt, stride=2, padding=3)
    assert_close_bf16(nchw(y.view(n, ho, wo, 64)), ref, "stem fwd")
    # wgrad through the patch GEMM
    dy = rnd(1, 1, n * ho * wo, 64, seed=5)
    dw = torch.zeros(64, 1, o.STEM_KPAD, device="cuda")
    o.conv_wgrad(o.conv_args(patches, dy, k=1, dw=dw))
    g = torch.empty(64, 3, 7, 7, device="cuda")
    o.unpack_wgrad(dw, g, 64, 3, 7, 3, o.STEM_KPAD)
    torch.cuda.synchronize()
    dyn = nchw(dy.view(n, ho, wo, 64))
    refw = torch.nn.grad.conv2d_weight(x, (64, 3, 7, 7), dyn, stride=2, padding=3)
    assert (g - refw).abs().max().item() <= 2e-3 * refw.abs().max().item() + 1e-4
    # dgrad: dpatches = dy @ W  (weights [k_pad][64] as the B operand), then col2im
    dpatch = torch.empty(1, 1, n * ho * wo, o.STEM_KPAD, dtype=torch.bfloat16, device="cuda")
    o.conv_fwd(o.conv_args(dy, dpatch, pk.w_t, k=1))
    dx = torch.empty(n, 3, h, w, device="cuda")
    o.stem_col2im(dpatch, dx)
    torch.cuda.synchronize()
    refx = torch.nn.grad.conv2d_input((n, 3, h, w), wt, dyn, stride=2, padding=3)
    assert (dx - refx).abs().max().item() <= 2e-2 * refx.abs().max().item()


def test_batchnorm_train_fwd_bwd():
    o = ops()
    n, h, w, c = 4, 16, 20, 64
    z = rnd(n, h, w, c, seed=1, scale=2.0) + 0.5
    res = rnd(n, h, w, c, seed=2)
    gamma = (torch.rand(c, device="cuda") + 0.5)
    beta = torch.randn(c, device="cuda") * 0.1
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    zf = nchw(z)
    stats = torch.zeros(o.STATS_REPLICAS, 2, c, device="cuda")
    stats[0, 0] = zf.sum((0, 2, 3))
    stats[0, 1] = (zf * zf).sum((0, 2, 3))
    mean, invstd, scale, shift = (torch.empty(c, device="cuda") for _ in range(4))
    o.bn_finalize(stats, n * h * w, gamma, beta, 1e-5, 0.1, rm, rv, mean, invstd, scale, shift)
    y = torch.empty_like(z)
    o.bn_apply(z, scale, shift, y, relu=True, res=res)
    torch.cuda.synchronize()
    zr = zf.clone().requires_grad_(True)
    g_ = gamma.clone().requires_grad_(True)
    b_ = beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    ref = F.relu(F.batch_norm(zr, rm2, rv2, g_, b_, training=True, momentum=0.1, eps=1e-5) + nchw(res))
    assert_close_bf16(nchw(y), ref.detach(), "bn_apply")
    assert torch.allclose(rm, rm2, atol=1e-5) and torch.allclose(rv, rv2, rtol=1e-4, atol=1e-5)
    dy = rnd(n, h, w, c, seed=3)
    (ref * nchw(dy)).sum().backward()
    sums = torch.zeros(2, c, device="cuda")
    o.bn_bwd_reduce(dy, y, z, mean, invstd, sums)
    dz, gout = torch.empty_like(z), torch.empty_like(z)
    dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    o.bn_bwd_apply(dy, y, z, mean, invstd, gamma, sums, dz, gout, dgamma, dbeta)
    # ReLU mask recomputed from z (no residual): must equal the y_relu path on a residual-free layer
    y2 = torch.empty_like(z)
    o.bn_apply(z, scale, shift, y2, relu=True)
    s_a, s_b = torch.zeros(2, c, device="cuda"), torch.zeros(2, c, device="cuda")
    o.bn_bwd_reduce(dy, y2, z, mean, invstd, s_a)
    o.bn_bwd_reduce(dy, None, z, mean, invstd, s_b, relu_scale=scale, relu_shift=shift)
    dz_a, dz_b = torch.empty_like(z), torch.empty_like(z)
    o.bn_bwd_apply(dy, y2, z, mean, invstd, gamma, s_a, dz_a)
    o.bn_bwd_apply(dy, None, z, mean, invstd, gamma, s_b, dz_b, relu_scale=scale, relu_shift=shift)
    torch.cuda.synchronize()
    assert torch.allclose(s_a, s_b, rtol=1e-4, atol=1e-3) and (dz_a.float() - dz_b.float()).abs().max().item() < 1e-2
    torch.cuda.synchronize()
    # mask differences at exactly-rounded-to-zero outputs are possible; compare with a tolerance on the sums
    assert torch.allclose(dgamma, g_.grad, rtol=2e-2, atol=2e-1), (dgamma - g_.grad).abs().max()
    assert torch.allclose(dbeta, b_.grad, rtol=2e-2, atol=2e-1)
    err = (nchw(dz) - zr.grad).abs()
    assert err.mean().item() < 5e-3 and (err > 0.05).float().mean().item() < 1e-3


def test_bn_small_channels():
    o = ops()
    for c in (16, 32, 512):
        n, h, w = 2, 8, 12
        z, dy = rnd(n, h, w, c, seed=1), rnd(n, h, w, c, seed=3)
        zf = nchw(z)
        mean, var = zf.mean((0, 2, 3)), zf.var((0, 2, 3), unbiased=False)
        invstd = (var + 1e-5).rsqrt()
        sums = torch.zeros(2, c, device="cuda")
        o.bn_bwd_reduce(dy, None, z, mean.contiguous(), invstd.contiguous(), sums)
        torch.cuda.synchronize()
        xh = (zf - mean[None, :, None, None]) * invstd[None, :, None, None]
        assert torch.allclose(sums[0], nchw(dy).sum((0, 2, 3)), rtol=1e-3, atol=1e-2)
        assert torch.allclose(sums[1], (nchw(dy) * xh).sum((0, 2, 3)), rtol=1e-3, atol=1e-2)


def test_maxpool_fwd_bwd():
    o = ops()
    n, h, w, c = 2, 16, 24, 64
    x = F.relu(rnd(n, h, w, c, seed=1))
    y = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device="cuda")
    o.maxpool_fwd(x, y)
    xr = nchw(x).requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    torch.cuda.synchronize()
    assert torch.equal(nchw(y), ref.detach())
    dy = rnd(n, h // 2, w // 2, c, seed=2)
    add = rnd(n, h, w, c, seed=3)
    dx = torch.empty_like(x)
    o.maxpool_bwd(x, y, dy, dx, add=add, relu_mask=True)
    torch.cuda.synchronize()
    (ref * nchw(dy)).sum().backward()
    want = (xr.grad + nchw(add)) * (nchw(x) > 0)
    assert_close_bf16(nchw(dx), want, "maxpool_bwd")
    # arg-max index path (what the engines use): forward stores the window position, backward never reads x / y
    idx = torch.full((n, h // 2, w // 2, c), 255, dtype=torch.uint8, device="cuda")
    y2 = torch.empty_like(y)
    o.maxpool_fwd(x, y2, idx=idx)
    dx2 = torch.empty_like(x)
    o.maxpool_bwd(x, y2, dy, dx2, add=add, idx=idx)
    torch.cuda.synchronize()
    assert torch.equal(y2, y) and int(idx.max()) <= 8
    assert_close_bf16(nchw(dx2), xr.grad + nchw(add), "maxpool_bwd(idx)")
    o.maxpool_fwd(x, y2, idx=idx, mask_nonpositive=True)      # ReLU mask of the pooled tensor's producer folded in
    o.maxpool_bwd(x, y2, dy, dx2, idx=idx)
    torch.cuda.synchronize()
    assert_close_bf16(nchw(dx2), xr.grad * (nchw(x) > 0), "maxpool_bwd(idx, masked)")


def test_upsample_and_fpn_add():
    o = ops()
    x = rnd(2, 8, 10, 32, seed=1)
    y = torch.empty(2, 16, 20, 32, dtype=torch.bfloat16, device="cuda")
    o.upsample2x_fwd(x, y)
    torch.cuda.synchronize()
    assert torch.equal(nchw(y), F.interpolate(nchw(x), scale_factor=2, mode="nearest"))
    dy = rnd(2, 16, 20, 32, seed=2)
    dx = torch.empty_like(x)
    o.upsample2x_bwd(dy, dx)
    torch.cuda.synchronize()
    assert_close_bf16(nchw(dx), F.avg_pool2d(nchw(dy), 2) * 4, "upsample bwd")
    for (hi, wi, ho, wo) in ((10, 10, 19, 19), (19, 19, 38, 38), (20, 20, 40, 40), (38, 38, 75, 75)):
        a = rnd(1, hi, wi, 16, seed=3)
        b = rnd(1, ho, wo, 16, seed=4)
        want = nchw(b) + F.interpolate(nchw(a), size=(ho, wo), mode="nearest")
        o.add_nearest_fwd(a, b)
        torch.cuda.synchronize()
        assert_close_bf16(nchw(b), want, "fpn add")
        g = rnd(1, ho, wo, 16, seed=5)
        ar = nchw(a).requires_grad_(True)
        (F.interpolate(ar, size=(ho, wo), mode="nearest") * nchw(g)).sum().backward()
        da = torch.empty_like(a)
        o.add_nearest_bwd(g, da)
        torch.cuda.synchronize()
        assert_close_bf16(nchw(da), ar.grad, "fpn add bwd")


def test_layout_converters_and_sigmoid_pack():
    o = ops()
    x = torch.randn(2, 256, 10, 12, device="cuda")
    y = torch.empty(2, 10, 12, 256, dtype=torch.bfloat16, device="cuda")
    o.nchw_f32_to_nhwc_bf16(x, y)
    torch.cuda.synchronize()
    assert torch.equal(y, nhwc(x))
    back = torch.empty_like(x)
    o.nhwc_bf16_to_nchw_f32(y, back)
    torch.cuda.synchronize()
    assert torch.equal(back, nchw(y))
    hal = torch.rand(2, 3, 8, 16, device="cuda")
    dhal = torch.randn(2, 3, 8, 16, device="cuda")
    dl = torch.full((2, 8, 16, 16), float("nan"), dtype=torch.bfloat16, device="cuda")
    db = torch.zeros(3, device="cuda")
    o.sigmoid_bwd_pack(dhal, hal, dl, db)
    torch.cuda.synchronize()
    want = dhal * hal * (1 - hal)
    assert torch.equal(dl[..., :3], nhwc(want)) and (dl[..., 3:] == 0).all()
    assert torch.allclose(db, want.sum((0, 2, 3)), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("hi,wi,s", [(512, 640, 640), (64, 96, 128), (1024, 1280, 300)])
def test_resize_nearest(hi, wi, s):
    o = ops()
    x = torch.rand(2, 3, hi, wi, device="cuda")
    y = torch.empty(2, 3, s, s, device="cuda")
    o.resize_nearest_fwd(x, y)
    torch.cuda.synchronize()
    xr = x.clone().requires_grad_(True)
    ref = F.interpolate(xr, size=[s, s])
    assert torch.equal(y, ref.detach())
    dy = torch.randn(2, 3, s, s, device="cuda")
    (ref * dy).sum().backward()
    dx = torch.empty_like(x)
    o.resize_nearest_bwd(dy, dx)
    torch.cuda.synchronize()
    assert torch.allclose(dx, xr.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind", ["mse", "l1"])
def test_regulariser(kind):
    o = ops()
    n, h, w = 2, 32, 48
    hal = torch.rand(n, 3, h, w, device="cuda").requires_grad_(True)
    rgb = torch.rand(n, 3, h, w, device="cuda")
    ir = torch.rand(n, 1, h, w, device="cuda")
    loss = torch.zeros(2, device="cuda")
    dhal = torch.empty(n, 3, h, w, device="cuda")
    o.regulariser(kind, hal.detach(), rgb, ir, 1.0, 0.5, loss, dhal)
    torch.cuda.synchronize()
    f = F.mse_loss if kind == "mse" else F.l1_loss
    l0, l1 = f(rgb, hal) * 1.0, f(ir.repeat(1, 3, 1, 1), hal) * 0.5
    (l0 + l1).backward()
    assert torch.allclose(loss, torch.stack([l0, l1]).detach(), rtol=1e-4)
    assert torch.allclose(dhal, hal.grad, rtol=1e-4, atol=1e-9)


def _nms_boxes(n, seed, clusters=40):
    g = torch.Generator().manual_seed(seed)
    centers = torch.rand(clusters, 2, generator=g) * 600
    c = centers[torch.randint(0, clusters, (n,), generator=g)] + torch.randn(n, 2, generator=g) * 12
    wh = torch.rand(n, 2, generator=g) * 80 + 4
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).cuda()
    scores = torch.rand(n, generator=g)
    if n:
        scores[::7] = scores[0]                               # ties: the stable sort decides
    return boxes, scores.cuda()


@pytest.mark.parametrize("n", [0, 1, 5, 63, 64, 65, 130, 1000, 4300, 8192])
@pytest.mark.parametrize("thr", [0.5, 0.7])
def test_nms_matches_torchvision(n, thr):
    """hd_nms vs torchvision.ops.nms (the op the reference's torchvision detectors call): identical indices, in order."""
    import torchvision
    o = ops()
    boxes, scores = _nms_boxes(n, seed=n + 1)
    ref = torchvision.ops.nms(boxes, scores, thr)
    got = o.nms(boxes, scores, thr)
    assert got.dtype == torch.int64 and torch.equal(got, ref)


def test_nms_batched_problems_and_coordinate_trick():
    import torchvision
    from hallucidet_b200 import detection as D
    o = ops()
    probs = [_nms_boxes(n, seed=100 + n) for n in (700, 0, 64, 3000, 129)]
    sorted_boxes, refs = [], []
    for b, s in probs:
        order = torch.sort(s, dim=0, descending=True, stable=True)[1] if b.shape[0] else torch.empty(0, dtype=torch.int64, device="cuda")
        sorted_boxes.append(b[order])
        refs.append((order, torchvision.ops.nms(b, s, 0.6)))
    keeps = o.nms_sorted_batch(sorted_boxes, 0.6)
    for (order, ref), keep in zip(refs, keeps):
        assert torch.equal(order.masked_select(keep), ref)
    b, s = _nms_boxes(2500, seed=9)
    idxs = torch.randint(0, 5, (2500,), device="cuda")
    assert torch.equal(D.batched_nms(b, s, idxs, 0.7), torchvision.ops.batched_nms(b, s, idxs, 0.7))


def test_nms_valid_mask_and_many_problems():
    """hd_nms_valid: boxes masked out in place (valid = 0) against torchvision.ops.nms on the compacted list, for 40 problems in
    one call (the proposal filter's (image, level) layout) -- a masked box is never kept and suppresses nothing."""
    import torchvision
    o = ops()
    sizes = [1000, 1000, 1000, 1000, 300] * 8
    g = torch.Generator().manual_seed(3)
    boxes_all, valid_all, refs, offs = [], [], [], [0]
    for i, n in enumerate(sizes):
        b, s = _nms_boxes(n, seed=500 + i)
        order = torch.sort(s, dim=0, descending=True, stable=True)[1]
        b, s = b[order], s[order]
        v = (torch.rand(n, generator=g) < (0.8 if i % 3 else 0.2)).cuda()
        live = torch.nonzero(v)[:, 0]
        ref = torch.zeros(n, dtype=torch.bool, device="cuda")
        ref[live[torchvision.ops.nms(b[live], s[live], 0.7)]] = True
        boxes_all.append(b); valid_all.append(v); refs.append(ref); offs.append(offs[-1] + n)
    keep = o.nms_sorted_flat(torch.cat(boxes_all).contiguous(), offs, 0.7, valid=torch.cat(valid_all).contiguous())
    assert torch.equal(keep, torch.cat(refs))


def test_multi_layer_pack_and_unpack_match_single_layer_calls():
    """hd_pack_conv_weights / hd_unpack_wgrads (one launch for every layer) against the per-layer entry points."""
    o = ops()
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 3, 7, 160, False), (64, 64, 3, None, True), (16, 32, 3, None, True), (256, 128, 1, None, True), (16, 16, 3, None, True),
              (128, 64, 3, None, True), (48, 32, 1, None, True), (32, 128, 3, None, True)]     # (tiled path: no padding, cout % 16 == 0)
    ws, singles, multis = [], [], []
    for cout, cin, k, k_pad, dg in shapes:
        w = torch.randn(cout, cin, k, k, generator=g).cuda()
        ws.append(w)
        singles.append(o.PackedConv(cout, cin, k, "cuda", need_dgrad=dg, need_t=not dg, k_pad=k_pad).pack(w))
        multis.append(o.PackedConv(cout, cin, k, "cuda", need_dgrad=dg, need_t=not dg, k_pad=k_pad))
        for t in (multis[-1].w_fwd, multis[-1].w_dgrad, multis[-1].w_t):
            if t is not None:
                t.fill_(float("nan"))
    o.pack_conv_weights(o.pack_table(multis, ws, "cuda"))
    torch.cuda.synchronize()
    for a, b in zip(singles, multis):
        for x, y in ((a.w_fwd, b.w_fwd), (a.w_dgrad, b.w_dgrad), (a.w_t, b.w_t)):
            assert (x is None) == (y is None)
            if x is not None:
                assert torch.equal(x, y)
    entries, refs = [], []
    for cout, cin, k, k_pad, dg in shapes:
        row = k_pad if k_pad else k * k * cin
        dw = torch.randn(cout, row, generator=g).cuda()
        ga, gb = torch.empty(cout, cin, k, k, device="cuda"), torch.full((cout, cin, k, k), float("nan"), device="cuda")
        o.unpack_wgrad(dw, ga, cout, cin, k, cin, row)
        entries.append((dw, gb, cout, cin, k, cin, row))
        refs.append(ga)
    o.unpack_wgrads(o.unpack_table(entries, "cuda"))
    torch.cuda.synchronize()
    assert all(torch.equal(r, e[1]) for r, e in zip(refs, entries))


@pytest.mark.parametrize("h,w,scale,c", [(160, 160, 0.25, 256), (40, 40, 1.0 / 16, 256), (20, 24, 1.0 / 32, 64)])
def test_roi_align_bwd_matches_torchvision(h, w, scale, c):
    """hd_roi_align_bwd_nhwc (+ layout conversion) against the autograd of torchvision.ops.roi_align (aligned=False,
    sampling_ratio 2, 7x7): RoIs of all sizes, partly outside the image, degenerate, larger than the image."""
    import torchvision
    o = ops()
    g = torch.Generator().manual_seed(h)
    n, k = 3, 200
    feat = torch.randn(n, c, h, w, generator=g).cuda().requires_grad_(True)
    img_w, img_h = w / scale, h / scale
    xy = torch.rand(k, 2, generator=g) * torch.tensor([img_w, img_h]) - 20
    wh = torch.rand(k, 2, generator=g) ** 2 * torch.tensor([img_w, img_h]) * 1.2 + 0.5
    wh[:5] = 0.0                                               # degenerate boxes (width / height clamp to one pixel)
    wh[5:8] = torch.tensor([img_w, img_h]) * 1.5               # larger than the image
    rois = torch.cat([torch.randint(0, n, (k, 1), generator=g).float(), xy, xy + wh], 1).cuda()
    out = torchvision.ops.roi_align(feat, rois, (7, 7), scale, 2, False)
    gout = torch.randn(out.shape, generator=g).cuda()
    (ref,) = torch.autograd.grad(out, feat, gout)
    got = o.roi_align_bwd(gout.contiguous(), rois, tuple(feat.shape), scale, 2)
    torch.cuda.synchronize()
    assert float(ref.abs().max()) > 0 and got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()))


@pytest.mark.parametrize("h,w,scale,c", [(160, 160, 0.25, 256), (40, 40, 1.0 / 16, 256), (20, 24, 1.0 / 32, 64)])
def test_roi_align_fwd_matches_torchvision(h, w, scale, c):
    """hd_roi_align_fwd_nhwc against torchvision.ops.roi_align: the same per-element expression, so the pooled features
    are expected to be bit-identical (reported if not; gated at 1e-6 relative)."""
    import torchvision
    o = ops()
    g = torch.Generator().manual_seed(h + 1)
    n, k = 3, 300
    feat = torch.randn(n, c, h, w, generator=g).cuda()
    img_w, img_h = w / scale, h / scale
    xy = torch.rand(k, 2, generator=g) * torch.tensor([img_w, img_h]) - 20
    wh = torch.rand(k, 2, generator=g) ** 2 * torch.tensor([img_w, img_h]) * 1.2 + 0.5
    wh[:5] = 0.0
    wh[5:8] = torch.tensor([img_w, img_h]) * 1.5
    rois = torch.cat([torch.randint(0, n, (k, 1), generator=g).float(), xy, xy + wh], 1).cuda()
    ref = torchvision.ops.roi_align(feat, rois, (7, 7), scale, 2, False)
    nhwc = o.nchw_to_nhwc_f32(feat)
    assert torch.equal(nhwc, feat.permute(0, 2, 3, 1).contiguous())
    got = o.roi_align_fwd(nhwc, rois, (7, 7), scale, 2)
    torch.cuda.synchronize()
    diff = (got - ref).abs().max().item()
    print(f"\n[roi_align fwd {h}x{w} c{c}] bit-identical: {torch.equal(got, ref)}, max |diff| {diff:.3e}")
    assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6 * float(ref.abs().max()))
    assert torch.equal(got, ref)         # holds with this toolchain: same expression, same contraction


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 24, 40, 64, 128), (8, 32, 40, 256, 256), (2, 64, 64, 16, 16), (2, 40, 64, 32, 32),
                                            (8, 128, 160, 64, 64), (2, 37, 41, 64, 64), (2, 48, 40, 128, 32),
                                            (4, 64, 80, 64, 64)])
def test_conv_fwd_fused_bn_finalize(n, h, w, cin, cout):
    """hd_conv_args.bn_fin: the last CTA of the convolution finalizes train-mode BatchNorm (mean / invstd / scale / shift and
    the running statistics, nn.BatchNorm2d semantics) from the per-CTA statistics rows -- against fp32 PyTorch on the
    bf16-rounded conv output, twice in a row (the per-layer counter must be left at zero), and bit-identical run to run."""
    o = ops()
    x = rnd(n, h, w, cin, seed=1)
    wt = (torch.randn(cout, cin, 3, 3, generator=torch.Generator().manual_seed(3)) / (cin * 9) ** 0.5).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(cout, cin, 3, "cuda").pack(wt)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    stats = torch.full((o.conv_fwd_tiles(x, 3, 1, cout=cout), 2, cout), float("nan"), device="cuda")
    gamma = torch.rand(cout, device="cuda") + 0.5
    beta = torch.randn(cout, device="cuda")
    rm, rv = torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda")
    mean, invstd, scale, shift = (torch.full((cout,), float("nan"), device="cuda") for _ in range(4))
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    fin = o.bn_fin(n * h * w, gamma, beta, 1e-5, 0.1, rm, rv, mean, invstd, scale, shift, counter)
    o.conv_fwd(o.conv_args(x, y, pk.w_fwd, k=3, stats=stats, bn_fin=fin))
    torch.cuda.synchronize()
    assert int(counter) == 0
    yq = nchw(y).double()
    m_ref = yq.mean((0, 2, 3))
    v_ref = yq.var((0, 2, 3), unbiased=False)
    is_ref = (v_ref + 1e-5).rsqrt()
    assert torch.allclose(mean.double(), m_ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(invstd.double(), is_ref, rtol=1e-4)
    assert torch.allclose(scale.double(), gamma.double() * is_ref, rtol=1e-4)
    assert torch.allclose(shift.double(), beta.double() - m_ref * gamma.double() * is_ref, rtol=1e-4, atol=1e-5)
    cnt = n * h * w
    assert torch.allclose(rm.double(), 0.1 * m_ref, rtol=1e-4, atol=1e-6)
    assert torch.allclose(rv.double(), 0.9 + 0.1 * v_ref * cnt / (cnt - 1), rtol=1e-4)
    first = [t.clone() for t in (mean, invstd, scale, shift, stats)]
    o.conv_fwd(o.conv_args(x, y, pk.w_fwd, k=3, stats=stats, bn_fin=fin))     # second launch: counter was reset, results identical
    torch.cuda.synchronize()
    assert int(counter) == 0
    for a, b in zip(first, (mean, invstd, scale, shift, stats)):
        assert torch.equal(a, b)
    assert torch.allclose(rm.double(), 0.19 * m_ref, rtol=1e-4, atol=1e-6)
    # the separate finalize kernel on the same rows gives the same coefficients
    mean2, invstd2, scale2, shift2 = (torch.empty(cout, device="cuda") for _ in range(4))
    o.bn_finalize(stats, cnt, gamma, beta, 1e-5, 0.1, None, None, mean2, invstd2, scale2, shift2)
    torch.cuda.synchronize()
    assert torch.allclose(scale2, scale, rtol=1e-6) and torch.allclose(shift2, shift, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,h,w,c,mode", [(4, 16, 20, 64, "yrelu"), (2, 8, 12, 16, "direct"), (2, 8, 12, 512, "plain"),
                                          (8, 64, 80, 128, "direct"),     # slice kept in shared memory (config-2 layer2 size)
                                          (8, 128, 160, 64, "yrelu"),     # re-read mode (config-2 layer1 size)
                                          (3, 50, 70, 32, "direct")])
def test_bn_bwd_fused_matches_two_pass_and_torch(n, h, w, c, mode):
    """hd_bn_bwd_fused (reduce + grid barrier + apply in one persistent kernel) against the two-launch path and autograd,
    for the three ReLU-mask sources; launched several times to exercise the self-resetting barrier."""
    o = ops()
    z = rnd(n, h, w, c, seed=1, scale=2.0) + 0.5
    dy = rnd(n, h, w, c, seed=3)
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda") * 0.1
    zf = nchw(z)
    mean = zf.mean((0, 2, 3)).contiguous()
    invstd = (zf.var((0, 2, 3), unbiased=False) + 1e-5).rsqrt().contiguous()
    scale = (gamma * invstd).contiguous()
    shift = (beta - mean * scale).contiguous()
    res = rnd(n, h, w, c, seed=2)
    y = torch.empty_like(z)
    o.bn_apply(z, scale, shift, y, relu=True, res=res if mode == "yrelu" else None)
    kw = {}
    yr = None
    if mode == "yrelu":
        yr = y
    elif mode == "direct":
        kw = dict(relu_scale=scale, relu_shift=shift)
    s_ref = torch.zeros(2, c, device="cuda")
    o.bn_bwd_reduce(dy, yr, z, mean, invstd, s_ref, **kw)
    dz_ref, g_ref = torch.empty_like(z), torch.empty_like(z)
    dg_ref, db_ref = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    o.bn_bwd_apply(dy, yr, z, mean, invstd, gamma, s_ref, dz_ref, g_ref, dg_ref, db_ref, **kw)
    barrier = torch.zeros(2, dtype=torch.int32, device="cuda")
    for rep in range(3):
        sums = torch.zeros(2, c, device="cuda")
        dz, gout = torch.full_like(z, float("nan")), torch.full_like(z, float("nan"))
        dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        o.bn_bwd_fused(dy, yr, z, mean, invstd, gamma, sums, dz, barrier, gout, dgamma, dbeta, **kw)
        torch.cuda.synchronize()
        assert int(barrier[0]) == 0 and int(barrier[1]) == rep + 1
        assert torch.equal(gout, g_ref)
        assert torch.allclose(sums, s_ref, rtol=1e-4, atol=1e-2 * max(1.0, s_ref.abs().max().item() * 1e-3))
        assert torch.allclose(dgamma, dg_ref, rtol=1e-4, atol=1e-2) and torch.allclose(dbeta, db_ref, rtol=1e-4, atol=1e-2)
        err = (dz.float() - dz_ref.float()).abs().max().item()
        assert err <= 2 ** -7 * dz_ref.float().abs().max().item() + 1e-6, err      # same formula; fp32 sums differ in the last bits
    if mode == "plain":                                                             # autograd cross-check (no mask)
        zr = zf.clone().requires_grad_(True)
        out = F.batch_norm(zr, None, None, gamma, beta, training=True, eps=1e-5)
        (out * nchw(dy)).sum().backward()
        assert_close_bf16(nchw(dz), zr.grad, "bn_bwd_fused vs autograd")


def test_conv_streamk_epilogue_stats_add_mask_deterministic():
    """Stream-K launches (partial accumulators through the workspace, reduced in CTA order): fused epilogue operands and the
    BatchNorm statistics are unaffected, results are bit-identical run to run, and HD-level parity holds vs fp32 PyTorch."""
    o = ops()
    from hallucidet_b200 import _lib
    if not _lib.load().hd_conv_has_streamk():
        pytest.skip("library built without the experimental stream-K paths (HD_BUILD_STREAMK=1)")
    was, o.STREAMK = o.STREAMK, True                      # opt-in path (HD_STREAMK=1): conv_args attaches the workspace
    try:
        _streamk_case(o)
    finally:
        o.STREAMK = was


def _streamk_case(o):
    n, h, w, cin, cout = 8, 32, 40, 256, 256
    x = rnd(n, h, w, cin, seed=1)
    wt = (torch.randn(cout, cin, 3, 3, generator=torch.Generator().manual_seed(3)) / (cin * 9) ** 0.5).to(torch.bfloat16).float().cuda()
    pk = o.PackedConv(cout, cin, 3, "cuda").pack(wt)
    ref = F.conv2d(nchw(x), wt, padding=1)
    # (a) statistics epilogue (staged store)
    outs = []
    for _ in range(2):
        y = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
        stats = torch.full((o.conv_fwd_tiles(x, 3, 1, cout=cout), 2, cout), float("nan"), device="cuda")
        o.conv_fwd(o.conv_args(x, y, pk.w_fwd, k=3, stats=stats))
        torch.cuda.synchronize()
        outs.append((y, stats))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert_close_bf16(nchw(outs[0][0]), ref, "stream-K fwd")
    yq = nchw(outs[0][0])
    ssum = outs[0][1].sum(0)
    assert torch.allclose(ssum[0], yq.sum((0, 2, 3)), rtol=2e-3, atol=1e-2)
    assert torch.allclose(ssum[1], (yq * yq).sum((0, 2, 3)), rtol=2e-3, atol=1e-2)
    # (b) register-store epilogue with bias + add + relu + mask
    bias = torch.randn(cout, device="cuda")
    add, mask = rnd(n, h, w, cout, seed=5), rnd(n, h, w, cout, seed=6)
    y2 = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    o.conv_fwd(o.conv_args(x, y2, pk.w_fwd, k=3, bias=bias, add=add, relu=True, mask=mask))
    torch.cuda.synchronize()
    ref2 = F.relu(ref + bias.view(1, -1, 1, 1) + nchw(add)) * (nchw(mask) > 0)
    assert_close_bf16(nchw(y2), ref2, "stream-K fused epilogue")
    ws = o.conv_workspace(x.device)
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0        # every partial flag was handed back


@pytest.mark.parametrize("n,h,w", [(2, 40, 40), (2, 5, 5), (2, 3, 3), (2, 75, 75)])
def test_head_tower_kernel_shapes(n, h, w):
    """The launches hallucidet_b200/heads.py issues: 1x1 256 -> 16 predictor with a channels-last fp32 output, its
    input gradient (K = 16 per tap) with a ReLU mask, and the 3x3 input gradient written as channels-last fp32 only."""
    o = ops()
    c = 256
    x = rnd(n, h, w, c, seed=1)
    # predictor forward
    wp = (torch.randn(16, c, 1, 1, generator=torch.Generator().manual_seed(3)) / c ** 0.5).to(torch.bfloat16).float().cuda()
    bias = torch.randn(16, device="cuda")
    pk = o.PackedConv(16, c, 1, "cuda").pack(wp)
    dummy = torch.empty(n, h, w, 16, dtype=torch.bfloat16, device="cuda")
    pred = torch.full((n, h, w, 16), float("nan"), device="cuda")
    o.conv_fwd(o.conv_args(x, dummy, pk.w_fwd, k=1, bias=bias, out_f32=pred, out_f32_channels=16, out_f32_nhwc=True, store_bf16=False))
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(x), wp, bias)
    assert torch.allclose(pred.permute(0, 3, 1, 2), ref, rtol=1e-3, atol=1e-3 * ref.abs().max().item())
    # predictor input gradient with the ReLU mask
    dy = rnd(n, h, w, 16, seed=4)
    mask = rnd(n, h, w, c, seed=5)
    g = torch.full((n, h, w, c), float("nan"), dtype=torch.bfloat16, device="cuda")
    o.conv_dgrad(o.conv_args(dy, g, pk.w_dgrad, k=1, mask=mask))
    torch.cuda.synchronize()
    ref_g = torch.nn.grad.conv2d_input((n, c, h, w), wp, nchw(dy)) * (nchw(mask) > 0)
    assert_close_bf16(nchw(g), ref_g, "predictor dgrad (K=16, mask)")
    # 3x3 input gradient, fp32 channels-last output only
    w3 = (torch.randn(c, c, 3, 3, generator=torch.Generator().manual_seed(6)) / (c * 9) ** 0.5).to(torch.bfloat16).float().cuda()
    pk3 = o.PackedConv(c, c, 3, "cuda").pack(w3)
    dy3 = rnd(n, h, w, c, seed=7)
    dx = torch.full((n, h, w, c), float("nan"), device="cuda")
    dummy3 = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    o.conv_dgrad(o.conv_args(dy3, dummy3, pk3.w_dgrad, k=3, out_f32=dx, out_f32_channels=c, out_f32_nhwc=True, store_bf16=False))
    torch.cuda.synchronize()
    ref_dx = torch.nn.grad.conv2d_input((n, c, h, w), w3, nchw(dy3), padding=1)
    assert torch.allclose(dx.permute(0, 3, 1, 2), ref_dx, rtol=1e-3, atol=2e-3 * ref_dx.abs().max().item())
    # 3x3 forward with bias + ReLU (tower conv) and its masked input gradient
    b3 = torch.randn(c, device="cuda")
    hid = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    o.conv_fwd(o.conv_args(x, hid, pk3.w_fwd, k=3, bias=b3, relu=True))
    g2 = torch.full((n, h, w, c), float("nan"), dtype=torch.bfloat16, device="cuda")
    o.conv_dgrad(o.conv_args(dy3, g2, pk3.w_dgrad, k=3, mask=hid))
    torch.cuda.synchronize()
    assert_close_bf16(nchw(hid), F.relu(F.conv2d(nchw(x), w3, b3, padding=1)), "tower conv")
    assert_close_bf16(nchw(g2), ref_dx * (nchw(hid) > 0), "tower dgrad (mask)")


@pytest.mark.parametrize("cin,dtype", [(3, "f32"), (1, "f32"), (1, "u8")])
@pytest.mark.parametrize("n,h,w", [(2, 64, 96), (1, 150, 300), (2, 128, 128)])
def test_stem_fwd_fused(cin, dtype, n, h, w):
    """hd_stem_fwd (halo-patch 7x7/2 stem, csrc/stem_conv.cu) against F.conv2d on the bf16-rounded input and weights:
    three-channel input, and the single replicated plane (fp32 or uint8 * 1/255) with the channel-summed filter; the BatchNorm
    statistics / fused finalize variant and the bias + ReLU variant."""
    o = ops()
    g = torch.Generator().manual_seed(7)
    wt = (torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5).cuda()
    if dtype == "u8":
        xu = torch.randint(0, 256, (n, 1, h, w), generator=g, dtype=torch.uint8).cuda()
        x_in, scale = xu, 1.0 / 255.0
        xf = (xu.float() * scale).to(torch.bfloat16).float()
    else:
        x_in = torch.rand(n, cin, h, w, generator=g).cuda()
        scale = 1.0
        xf = x_in.to(torch.bfloat16).float()
    if cin == 1:
        w_eff = wt.sum(1, keepdim=True).to(torch.bfloat16).float()
        ref = F.conv2d(xf, w_eff, stride=2, padding=3)
    else:
        ref = F.conv2d(xf, wt.to(torch.bfloat16).float(), stride=2, padding=3)
    ho, wo = h // 2, w // 2
    # (a) raw output + statistics + fused finalize
    y = torch.full((n, ho, wo, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    stats = torch.full((o.stem_fwd_rows(x_in), 2, 64), float("nan"), device="cuda")
    gamma, beta = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda")
    rm, rv = torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
    mean, invstd, sc, sh = (torch.full((64,), float("nan"), device="cuda") for _ in range(4))
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    fin = o.bn_fin(n * ho * wo, gamma, beta, 1e-5, 0.1, rm, rv, mean, invstd, sc, sh, counter)
    o.stem_fwd(x_in, wt, y, x_scale=scale, stats=stats, bn_fin=fin)
    torch.cuda.synchronize()
    assert_close_bf16(nchw(y), ref, "stem fwd")
    yq = nchw(y).double()
    ssum = stats.double().sum(0)
    assert torch.allclose(ssum[0], yq.sum((0, 2, 3)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(ssum[1], (yq * yq).sum((0, 2, 3)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(mean.double(), yq.mean((0, 2, 3)), rtol=1e-4, atol=1e-5)
    assert torch.allclose(invstd.double(), (yq.var((0, 2, 3), unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)
    assert int(counter) == 0
    # (b) folded scale + bias + ReLU (frozen backbone / eval-mode U-Net)
    wsc, bias = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1
    y2 = torch.full((n, ho, wo, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    o.stem_fwd(x_in, wt, y2, x_scale=scale, w_scale=wsc, bias=bias, relu=True)
    torch.cuda.synchronize()
    if cin == 1:
        w2 = (wt.sum(1, keepdim=True) * wsc.view(-1, 1, 1, 1)).to(torch.bfloat16).float()
    else:
        w2 = (wt * wsc.view(-1, 1, 1, 1)).to(torch.bfloat16).float()
    ref2 = F.relu(F.conv2d(xf, w2, bias, stride=2, padding=3))
    assert_close_bf16(nchw(y2), ref2, "stem fwd bias relu")
