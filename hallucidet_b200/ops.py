"""Tensor-level wrappers over the C ABI (include/hallucidet_b200.h).

Activations are ``torch.bfloat16`` tensors of shape ``[N, H, W, C]`` (NHWC, contiguous, C % 16 == 0);
module-edge tensors are ``torch.float32`` NCHW.  All launches go to ``torch.cuda.current_stream()``.
Nothing here computes on the CPU; PyTorch operators only appear around the NMS / RoIAlign calls (sort, gather, allocation),
exactly as in the torchvision ops they replace.  A missing library / wrong device raises.
"""
import ctypes

import torch

from . import _lib
from ._lib import HdRoiGatherArgs, HdAct, HdBnFin, HdConvArgs, check

STATS_REPLICAS = 16      # legacy name: default row count for small test problems (see conv_fwd_tiles)
STEM_KPAD = 160          # 7*7*3 = 147 -> 5 k-blocks of 32
# Stream-K (hd_conv_args.workspace) is opt-in (HD_STREAMK=1): measured on the config-2 deep layers it shortens the main loop
# (12.5 -> 8.7 us at 256 ch, 32x40) but the fp32 partial tiles through L2 cost more than that (profiles/r2_conv_timeline_v1.txt)
STREAMK = __import__("os").environ.get("HD_STREAMK", "0") == "1"

LAUNCHES = 0             # kernels launched through this module (one per C-ABI compute call; nms / roi_align_bwd add their second)
PROFILE = None           # set to a list to record (name, algorithmic FLOPs, start event, end event, shape, bytes) per launch
RECORD = None            # set to a list to record the signature of every conv launch (launch_signature) -- parity tests replay them


class _Timed:
    """Context manager that brackets one launch with CUDA events on the launching stream when PROFILE is on."""

    def __init__(self, name, flops=0.0, desc=""):
        self.name, self.flops, self.desc = name, flops, desc

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        global LAUNCHES
        LAUNCHES += 1
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.name, self.flops, self.e0, e1, self.desc, getattr(self, "bytes", 0.0)))
        return False


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def act(t):
    if t is None:
        return HdAct(None, 0, 0, 0, 0)
    assert t.dtype == torch.bfloat16 and t.dim() == 4 and t.is_contiguous(), (t.dtype, t.shape, t.is_contiguous())
    return HdAct(t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.shape[3])


def round_up(a, b):
    return (a + b - 1) // b * b


def conv_args(x0, y0, w=None, k=3, stride=1, x1=None, y1=None, bias=None, add=None, mask=None, relu=False, sigmoid=False,
              stats=None, out_f32=None, out_f32_channels=0, store_bf16=True, phase_mask=0, dw=None, split_k=0, out_f32_nhwc=False,
              algo_cin=None, algo_cout=None, bn_fin=None):
    a = HdConvArgs()
    if bn_fin is not None:
        a.bn_fin = ctypes.pointer(bn_fin)
        a._bn_fin_keepalive = bn_fin
    a.algo = (algo_cin, algo_cout)
    a.x0, a.x1, a.y0, a.y1 = act(x0), act(x1), act(y0), act(y1)
    if w is not None and STREAMK and x0.is_cuda:
        ws = conv_workspace(x0.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.w = w.data_ptr() if w is not None else None
    a.kh = a.kw = k
    a.stride = stride
    a.bias = bias.data_ptr() if bias is not None else None
    a.add = add.data_ptr() if add is not None else None
    a.mask = mask.data_ptr() if mask is not None else None
    a.relu, a.sigmoid = int(relu), int(sigmoid)
    a.stats = stats.data_ptr() if stats is not None else None
    a.stats_replicas = stats.shape[0] if stats is not None else 0
    a.out_f32_nchw = out_f32.data_ptr() if out_f32 is not None else None
    a.out_f32_channels = out_f32_channels
    a.out_f32_nhwc = int(bool(out_f32_nhwc))
    a.store_bf16 = int(store_bf16)
    a.phase_mask = phase_mask
    a.dw = dw.data_ptr() if dw is not None else None
    a.split_k = split_k
    return a


def _conv_flops(a, kind):
    """Algorithmic FLOPs (2*MACs) of the convolution the launch implements (padding channels excluded)."""
    k2 = a.kh * a.kw
    if kind == "fwd":
        cin, cout, pix = a.x0.c + a.x1.c, a.y0.c, a.y0.n * a.y0.h * a.y0.w
    elif kind == "dgrad":       # x0 = dY (conv output), y = dX (conv input)
        cin, cout, pix = a.y0.c + a.y1.c, a.x0.c, a.x0.n * a.x0.h * a.x0.w
    else:                       # wgrad: x = conv input, y0 = dY
        cin, cout, pix = a.x0.c + a.x1.c, a.y0.c, a.y0.n * a.y0.h * a.y0.w
    algo = getattr(a, "algo", (None, None))
    if algo[0] is not None:
        cin, k2 = algo[0], 1
    if algo[1] is not None:
        cout = algo[1]
    return 2.0 * pix * cin * cout * k2


def _is_narrow(a, kind):
    """Mirror of narrow_conv_eligible / narrow_wgrad_eligible (csrc/narrow_conv.cu): which kernel a launch runs on.
    Only used to label PROFILE records."""
    if a.kh != 3 or a.stride != 1 or a.x1.c or a.x0.c not in (16, 32) or a.y0.c not in (16, 32):
        return False
    if kind == "wgrad":
        return True
    if a.y1.c or a.add or a.mask or (kind == "dgrad" and a.stats) or (a.stats and a.out_f32_nchw):
        return False
    return True


def _conv_bytes(a, kind):
    """Algorithmic HBM bytes of a launch: every operand element once (bf16 activations, fp32 side outputs / gradients)."""
    px = a.x0.n * a.x0.h * a.x0.w
    py = a.y0.n * a.y0.h * a.y0.w
    b = 2.0 * px * (a.x0.c + a.x1.c)
    if kind == "wgrad":
        return b + 2.0 * py * a.y0.c + 4.0 * a.kh * a.kw * (a.x0.c + a.x1.c) * a.y0.c
    if a.store_bf16:
        b += 2.0 * py * (a.y0.c + a.y1.c)
    if a.out_f32_nchw:
        b += 4.0 * py * a.out_f32_channels
    if a.add:
        b += 2.0 * py * a.y0.c
    if a.mask:
        b += 2.0 * py * a.y0.c
    return b + 2.0 * a.kh * a.kw * (a.x0.c + a.x1.c) * (a.y0.c + a.y1.c)


def launch_signature(a, kind):
    """Everything that selects a code path inside hd_conv_fwd / hd_conv_dgrad / hd_conv_wgrad for one launch: operand shapes,
    filter, stride and which fused epilogue operands are present (hashable; the config-size parity tests replay each distinct one)."""
    sh = lambda t: (t.n, t.h, t.w, t.c)
    return (kind, sh(a.x0), a.x1.c, sh(a.y0), a.y1.c, a.kh, a.stride, bool(a.bias), bool(a.add), bool(a.mask), int(a.relu),
            int(a.sigmoid), bool(a.stats), bool(a.out_f32_nchw), int(a.out_f32_channels), int(a.out_f32_nhwc), int(a.store_bf16),
            int(a.phase_mask), bool(a.add) and a.add == a.y0.ptr)


def _conv_timed(a, kind):
    if RECORD is not None:
        RECORD.append(launch_signature(a, kind))
    if PROFILE is None:
        return _Timed("conv_" + kind)
    t = _Timed("conv_" + kind + ("_narrow" if _is_narrow(a, kind) else ""), _conv_flops(a, kind), _conv_desc(a))
    t.bytes = _conv_bytes(a, kind)
    return t


def _conv_desc(a):
    return (f"x0[{a.x0.n},{a.x0.h},{a.x0.w},{a.x0.c}] x1c{a.x1.c} y0[{a.y0.n},{a.y0.h},{a.y0.w},{a.y0.c}] y1c{a.y1.c} "
            f"k{a.kh} s{a.stride}")


_WORKSPACES = {}


def conv_workspace(device):
    """The stream-K scratch of hd_conv_fwd / hd_conv_dgrad (hd_conv_args.workspace): one zero-initialised buffer per device,
    shared by all convolution launches -- they are all issued on one stream (weight gradients, the only side-stream
    convolutions, do not use it)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _WORKSPACES.get(key)
    if ws is None:
        ws = _WORKSPACES[key] = torch.zeros(int(_lib.load().hd_conv_workspace_bytes()), dtype=torch.uint8, device=device)
    return ws


def conv_fwd(args):
    with _conv_timed(args, "fwd"):
        check(_lib.load().hd_conv_fwd(ctypes.byref(args), _stream()), "hd_conv_fwd")


def conv_fwd_tiles(x0, k=3, stride=1, cout=None):
    """Rows the partial BN statistics buffer needs for a forward conv over x0 (host-only query).  With ``cout`` the
    library can tell which kernel will run (the 16/32-channel layers write one row per CTA instead of one per tile)."""
    a = HdConvArgs()
    a.x0 = act(x0)
    a.kh = a.kw = k
    a.stride = stride
    if cout is not None:
        a.y0.c = int(cout)
    n = _lib.load().hd_conv_fwd_tiles(ctypes.byref(a))
    if n <= 0:
        raise RuntimeError(f"hd_conv_fwd_tiles failed ({n})")
    return n


def conv_dgrad(args):
    with _conv_timed(args, "dgrad"):
        check(_lib.load().hd_conv_dgrad(ctypes.byref(args), _stream()), "hd_conv_dgrad")


def conv_wgrad(args):
    with _conv_timed(args, "wgrad"):
        check(_lib.load().hd_conv_wgrad(ctypes.byref(args), _stream()), "hd_conv_wgrad")


class PackedConv:
    """bf16 GEMM operands of one conv layer, packed from the fp32 OIHW master weight."""

    def __init__(self, cout, cin, k, device, need_dgrad=True, need_t=False, k_pad=None):
        self.cout, self.cin, self.k = cout, cin, k
        self.cout_pad = round_up(cout, 16)
        self.cin_pad = round_up(cin, 16)
        self.k_pad = k_pad if k_pad is not None else k * k * cin
        self.w_fwd = torch.empty(self.cout_pad, self.k_pad, dtype=torch.bfloat16, device=device)
        self.w_dgrad = torch.empty(self.cin_pad, k * k * cout, dtype=torch.bfloat16, device=device) if need_dgrad else None
        self.w_t = torch.empty(self.k_pad, self.cout_pad, dtype=torch.bfloat16, device=device) if need_t else None

    def pack(self, w_oihw, scale=None):
        assert w_oihw.dtype == torch.float32 and w_oihw.is_contiguous() and tuple(w_oihw.shape) == (self.cout, self.cin, self.k, self.k)
        with _Timed("pack_conv_weight"):
            check(_lib.load().hd_pack_conv_weight(_ptr(w_oihw), _ptr(scale), self.cout, self.cin, self.k, self.k, _ptr(self.w_fwd),
                                                  self.cout_pad, self.k_pad, _ptr(self.w_dgrad), self.cin_pad, _ptr(self.w_t),
                                                  _stream()), "hd_pack_conv_weight")
        return self


def unpack_wgrad(dw_packed, grad_oihw, cout, cin, k, tap_stride, row_stride, scale=1.0):
    with _Timed("unpack_wgrad"):
        check(_lib.load().hd_unpack_wgrad(_ptr(dw_packed), _ptr(grad_oihw), cout, cin, k, k, tap_stride, row_stride, scale, _stream()),
              "hd_unpack_wgrad")


class _DescTable:
    """Device-resident descriptor table for the whole-network pack / unpack launches."""

    def __init__(self, descs, device):
        arr = (type(descs[0]) * len(descs))(*descs)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
        self.dev = raw.to(device)
        self.n = len(descs)
        self.total_blocks = descs[-1].first_block + descs[-1]._blocks
        self.keys = [d._key for d in descs]


def multi_blocks(elements):
    return int(_lib.load().hd_multi_blocks(ctypes.c_int64(int(elements))))


def pack_tiled_ok(pk):
    """Mirror of pack_tiled_ok (csrc/elementwise.cu): layers whose master weight is read exactly once by the pack launch -- the
    ones hd_adam_pack_conv_weights can update in place."""
    return (pk.w_dgrad is not None and pk.w_t is None and pk.k * pk.k <= 9 and pk.cout % 16 == 0 and pk.cin % 16 == 0
            and pk.cout_pad == pk.cout and pk.cin_pad == pk.cin and pk.k_pad == pk.k * pk.k * pk.cin)


def pack_table(packed_convs, weights, device, adam=None):
    """Descriptor table packing every (PackedConv, fp32 OIHW weight) pair in one launch (training mode: no BN scale).
    ``adam``: optional list of (grad, exp_avg, exp_avg_sq) per layer (None entries = pack only) for adam_pack_conv_weights."""
    descs, first = [], 0
    for i, (pk, w) in enumerate(zip(packed_convs, weights)):
        assert w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape) == (pk.cout, pk.cin, pk.k, pk.k)
        gmv = adam[i] if adam is not None else None
        if gmv is not None:
            assert pack_tiled_ok(pk) and all(t.dtype == torch.float32 and t.is_contiguous() and t.numel() == w.numel() for t in gmv)
        d = _lib.HdPackDesc(w.data_ptr(), None, pk.w_fwd.data_ptr(), pk.w_dgrad.data_ptr() if pk.w_dgrad is not None else None,
                            pk.w_t.data_ptr() if pk.w_t is not None else None, pk.cout, pk.cin, pk.k, pk.k, pk.cout_pad, pk.k_pad,
                            pk.cin_pad, first, *([t.data_ptr() for t in gmv] if gmv is not None else [None, None, None]))
        d._blocks = int(_lib.load().hd_pack_blocks(ctypes.byref(d)))
        d._key = w.data_ptr()
        first += d._blocks
        descs.append(d)
    return _DescTable(descs, device)


def pack_conv_weights(table):
    with _Timed("pack_conv_weights"):
        check(_lib.load().hd_pack_conv_weights(_ptr(table.dev), table.n, table.total_blocks, _stream()), "hd_pack_conv_weights")


def adam_args(lr, beta1, beta2, eps, step, grad_scale=1.0, clip=0.0):
    a = _lib.HdAdamArgs()
    a.lr, a.beta1, a.beta2, a.eps = float(lr), float(beta1), float(beta2), float(eps)
    a.bias_correction1, a.bias_correction2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    a.grad_scale, a.clip = float(grad_scale), float(clip)
    a.one_minus_beta1, a.one_minus_beta2 = 1.0 - beta1, 1.0 - beta2
    return a


def adam_pack_conv_weights(table, args):
    """clip + Adam + bf16 re-pack of every tiled layer in one pass over its master weight (hd_adam_pack_conv_weights)."""
    with _Timed("adam_pack_conv_weights"):
        check(_lib.load().hd_adam_pack_conv_weights(_ptr(table.dev), table.n, table.total_blocks, ctypes.byref(args), _stream()),
              "hd_adam_pack_conv_weights")


def adam_table(entries, device):
    """entries: (param, grad, exp_avg, exp_avg_sq) fp32 contiguous tensors -> descriptor table of hd_adam_multi."""
    descs, first = [], 0
    for p, g, m, v in entries:
        assert all(t.dtype == torch.float32 and t.is_contiguous() and t.numel() == p.numel() for t in (p, g, m, v))
        d = _lib.HdAdamDesc(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), first)
        d._blocks = multi_blocks(p.numel())
        d._key = p.data_ptr()
        first += d._blocks
        descs.append(d)
    return _DescTable(descs, device)


def adam_multi(table, args):
    with _Timed("adam_multi"):
        check(_lib.load().hd_adam_multi(_ptr(table.dev), table.n, table.total_blocks, ctypes.byref(args), _stream()), "hd_adam_multi")


def unpack_table(entries, device):
    """entries: (dw_packed, grad_oihw, cout, cin, k, tap_stride, row_stride) per layer."""
    descs, first = [], 0
    for dw, g, cout, cin, k, tap_stride, row_stride in entries:
        d = _lib.HdUnpackDesc(dw.data_ptr(), g.data_ptr(), cout, cin, k * k, tap_stride, row_stride, first, 1.0, 0)
        d._blocks = int(_lib.load().hd_unpack_blocks(cout, cin, k * k))     # one block = whole output-channel rows
        d._key = g.data_ptr()
        first += d._blocks
        descs.append(d)
    return _DescTable(descs, device)


def unpack_wgrads(table):
    with _Timed("unpack_wgrads"):
        check(_lib.load().hd_unpack_wgrads(_ptr(table.dev), table.n, table.total_blocks, _stream()), "hd_unpack_wgrads")


def stem_im2col(x_nchw, patches, k_pad=STEM_KPAD):
    n, c, h, w = x_nchw.shape
    assert c == 3 and x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    with _Timed("stem_im2col"):
        check(_lib.load().hd_stem_im2col(_ptr(x_nchw), _ptr(patches), n, h, w, k_pad, _stream()), "hd_stem_im2col")


STEM1_KPAD = 64          # single-channel stem: 7*7 = 49 -> one k-block of 64


def stem_im2col_1ch(x, patches, scale=1.0, k_pad=STEM1_KPAD):
    """x: [n, 1, h, w] float32 or uint8 (the IR plane; ``scale`` = 1/255 for camera bytes) -> bf16 patches [n*ho*wo][k_pad]."""
    n, c, h, w = x.shape
    assert c == 1 and x.is_contiguous() and x.dtype in (torch.float32, torch.uint8)
    with _Timed("stem_im2col"):
        check(_lib.load().hd_stem_im2col_1ch(_ptr(x), 0 if x.dtype == torch.float32 else 1, float(scale), _ptr(patches), n, h, w, k_pad,
                                             _stream()), "hd_stem_im2col_1ch")


def stem_fwd_rows(x):
    n, _, h, w = x.shape
    return int(_lib.load().hd_stem_fwd_rows(n, h, w))


def stem_fwd(x, w_oihw, y, x_scale=1.0, w_scale=None, bias=None, relu=False, stats=None, bn_fin=None):
    """Fused 7x7/2 stem (hd_stem_fwd): x [n, 3, h, w] fp32, or the single plane [n, 1, h, w] (fp32 / uint8) of a replicated input;
    w_oihw: the fp32 master weight [64, 3, 7, 7]; y: bf16 [n, h/2, w/2, 64]."""
    n, cin, h, w = x.shape
    assert cin in (1, 3) and x.is_contiguous() and x.dtype in (torch.float32, torch.uint8) and (cin == 1 or x.dtype == torch.float32)
    assert w_oihw.dtype == torch.float32 and w_oihw.is_contiguous() and tuple(w_oihw.shape) == (64, 3, 7, 7)
    assert y.dtype == torch.bfloat16 and y.is_contiguous() and tuple(y.shape) == (n, h // 2, w // 2, 64)
    pix = n * (h // 2) * (w // 2)
    with _Timed("stem_fwd", 2.0 * pix * 64 * 49 * cin, f"stem7x7 cin{cin} [{n},{h},{w}]") as t:
        t.bytes = float(x.numel() * x.element_size() + y.numel() * 2)
        check(_lib.load().hd_stem_fwd(_ptr(x), 0 if x.dtype == torch.float32 else 1, float(x_scale), cin, _ptr(w_oihw), _ptr(w_scale), _ptr(bias),
                                      int(relu), _ptr(y), n, h, w, _ptr(stats), stats.shape[0] if stats is not None else 0,
                                      ctypes.pointer(bn_fin) if bn_fin is not None else None, _stream()), "hd_stem_fwd")


def stem_col2im(dpatches, dx_nchw, k_pad=STEM_KPAD):
    n, c, h, w = dx_nchw.shape
    assert c == 3 and dx_nchw.dtype == torch.float32 and dx_nchw.is_contiguous()
    with _Timed("stem_col2im"):
        check(_lib.load().hd_stem_col2im(_ptr(dpatches), _ptr(dx_nchw), n, h, w, k_pad, _stream()), "hd_stem_col2im")


def bn_fin(count, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, scale, shift, counter):
    """hd_bn_fin for conv_args(bn_fin=...): the BatchNorm finalize runs in the tail of the convolution that produces the
    statistics (last CTA), instead of a separate hd_bn_finalize launch.  ``counter``: one zeroed int32 device word per layer.
    The struct holds raw pointers: the caller keeps the tensors alive."""
    assert counter.dtype == torch.int32 and counter.numel() == 1 and counter.is_cuda
    f = HdBnFin()
    f.count = float(count)
    f.gamma, f.beta = gamma.data_ptr(), beta.data_ptr()
    f.running_mean = running_mean.data_ptr() if running_mean is not None else None
    f.running_var = running_var.data_ptr() if running_var is not None else None
    f.mean, f.invstd, f.scale, f.shift = mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), shift.data_ptr()
    f.counter = counter.data_ptr()
    f.eps, f.momentum = float(eps), float(momentum)
    f._key = (gamma.data_ptr(), beta.data_ptr(), f.running_mean, f.running_var)
    return f


def bn_finalize(stats, count, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, scale, shift):
    c = gamma.numel()
    with _Timed("bn_finalize"):
        check(_lib.load().hd_bn_finalize(_ptr(stats), stats.shape[0], c, float(count), _ptr(gamma), _ptr(beta), eps, momentum,
                                         _ptr(running_mean), _ptr(running_var), _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift),
                                         _stream()), "hd_bn_finalize")


def bn_apply(z, scale, shift, y, relu=True, res=None, res_scale=None, res_shift=None):
    c = z.shape[-1]
    with _Timed("bn_apply"):
        check(_lib.load().hd_bn_apply(_ptr(z), _ptr(scale), _ptr(shift), _ptr(res), _ptr(res_scale), _ptr(res_shift), int(relu),
                                      _ptr(y), z.numel() // c, c, _stream()), "hd_bn_apply")


def bn_bwd_reduce(dy, y_relu, z, mean, invstd, sums, relu_scale=None, relu_shift=None):
    c = z.shape[-1]
    with _Timed("bn_bwd_reduce"):
        check(_lib.load().hd_bn_bwd_reduce(_ptr(dy), _ptr(y_relu), _ptr(relu_scale), _ptr(relu_shift), _ptr(z), _ptr(mean), _ptr(invstd), _ptr(sums), z.numel() // c, c,
                                           _stream()), "hd_bn_bwd_reduce")


def bn_bwd_apply(dy, y_relu, z, mean, invstd, gamma, sums, dz, g_out=None, dgamma=None, dbeta=None, relu_scale=None, relu_shift=None):
    c = z.shape[-1]
    n_pix = z.numel() // c
    with _Timed("bn_bwd_apply"):
        check(_lib.load().hd_bn_bwd_apply(_ptr(dy), _ptr(y_relu), _ptr(relu_scale), _ptr(relu_shift), _ptr(z), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(sums),
                                          float(n_pix), _ptr(dz), _ptr(g_out), _ptr(dgamma), _ptr(dbeta), n_pix, c, _stream()),
              "hd_bn_bwd_apply")


def bn_bwd_fused(dy, y_relu, z, mean, invstd, gamma, sums, dz, barrier, g_out=None, dgamma=None, dbeta=None, relu_scale=None, relu_shift=None):
    """bn_bwd_reduce + bn_bwd_apply as one persistent launch (hd_bn_bwd_fused).  ``sums`` zeroed by the caller; ``barrier``:
    int32[2] device tensor, zero-initialised once and then left to the kernel."""
    c = z.shape[-1]
    n_pix = z.numel() // c
    assert barrier.dtype == torch.int32 and barrier.numel() >= 2
    with _Timed("bn_bwd_fused"):
        check(_lib.load().hd_bn_bwd_fused(_ptr(dy), _ptr(y_relu), _ptr(relu_scale), _ptr(relu_shift), _ptr(z), _ptr(mean), _ptr(invstd),
                                          _ptr(gamma), _ptr(sums), float(n_pix), _ptr(dz), _ptr(g_out), _ptr(dgamma), _ptr(dbeta),
                                          n_pix, c, _ptr(barrier), _stream()), "hd_bn_bwd_fused")


def maxpool_fwd(x, y, idx=None, mask_nonpositive=False):
    """idx: optional uint8 [N, H/2, W/2, C] receiving the arg-max window position of every output element (for
    maxpool_bwd(idx=...)); mask_nonpositive folds a following ReLU mask into it."""
    ax, ay = act(x), act(y)
    assert idx is None or (idx.dtype == torch.uint8 and idx.is_contiguous() and tuple(idx.shape) == tuple(y.shape))
    with _Timed("maxpool_fwd"):
        check(_lib.load().hd_maxpool_fwd(ctypes.byref(ax), ctypes.byref(ay), _ptr(idx), int(mask_nonpositive), _stream()), "hd_maxpool_fwd")


def maxpool_bwd(x, y, dy, dx, add=None, relu_mask=False, idx=None):
    ax, ay = act(x), act(y)
    assert idx is None or not relu_mask, "with idx the ReLU mask is folded in by maxpool_fwd(mask_nonpositive=True)"
    with _Timed("maxpool_bwd"):
        check(_lib.load().hd_maxpool_bwd(ctypes.byref(ax), ctypes.byref(ay), _ptr(dy), _ptr(add), _ptr(dx), int(relu_mask), _ptr(idx),
                                         _stream()), "hd_maxpool_bwd")


def upsample2x_fwd(x, y):
    ax, ay = act(x), act(y)
    with _Timed("upsample2x_fwd"):
        check(_lib.load().hd_upsample2x_fwd(ctypes.byref(ax), ctypes.byref(ay), _stream()), "hd_upsample2x_fwd")


def upsample2x_bwd(dy, dx):
    ay, ax = act(dy), act(dx)
    with _Timed("upsample2x_bwd"):
        check(_lib.load().hd_upsample2x_bwd(ctypes.byref(ay), ctypes.byref(ax), _stream()), "hd_upsample2x_bwd")


def add_nearest_fwd(x, y):
    ax, ay = act(x), act(y)
    with _Timed("add_nearest_fwd"):
        check(_lib.load().hd_add_nearest_fwd(ctypes.byref(ax), ctypes.byref(ay), _stream()), "hd_add_nearest_fwd")


def add_nearest_bwd(dy, dx, accumulate=False):
    ay, ax = act(dy), act(dx)
    with _Timed("add_nearest_bwd"):
        check(_lib.load().hd_add_nearest_bwd(ctypes.byref(ay), ctypes.byref(ax), int(accumulate), _stream()), "hd_add_nearest_bwd")


def pad_hw(x, y):
    ax, ay = act(x), act(y)
    with _Timed("pad_hw"):
        check(_lib.load().hd_pad_hw(ctypes.byref(ax), ctypes.byref(ay), _stream()), "hd_pad_hw")


def crop_add_mask(dxp, dx, add=None, mask=None):
    ap, ax = act(dxp), act(dx)
    with _Timed("crop_add_mask"):
        check(_lib.load().hd_crop_add_mask(ctypes.byref(ap), _ptr(add), _ptr(mask), ctypes.byref(ax), _stream()), "hd_crop_add_mask")


def nchw_f32_to_nhwc_bf16(x, y, accumulate=False):
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[0] == y.shape[0] and x.shape[2:] == y.shape[1:3]
    ay = act(y)
    with _Timed("nchw_f32_to_nhwc_bf16"):
        check(_lib.load().hd_nchw_f32_to_nhwc_bf16(_ptr(x), ctypes.byref(ay), x.shape[1], int(accumulate), _stream()),
              "hd_nchw_f32_to_nhwc_bf16")


def nhwc_bf16_to_nchw_f32(x, y):
    assert y.dtype == torch.float32 and y.is_contiguous()
    ax = act(x)
    with _Timed("nhwc_bf16_to_nchw_f32"):
        check(_lib.load().hd_nhwc_bf16_to_nchw_f32(ctypes.byref(ax), _ptr(y), y.shape[1], _stream()), "hd_nhwc_bf16_to_nchw_f32")


def sigmoid_bwd_pack(dhal, hal, dlogits, dbias=None):
    ad = act(dlogits)
    with _Timed("sigmoid_bwd_pack"):
        check(_lib.load().hd_sigmoid_bwd_pack(_ptr(dhal), _ptr(hal), ctypes.byref(ad), hal.shape[1], _ptr(dbias), _stream()),
              "hd_sigmoid_bwd_pack")


def resize_nearest_fwd(x, y, mean=None, std=None):
    n, c, hi, wi = x.shape
    ho, wo = y.shape[-2:]
    with _Timed("resize_nearest_fwd"):
        check(_lib.load().hd_resize_nearest_fwd(_ptr(x), _ptr(y), n, c, hi, wi, ho, wo, _ptr(mean), _ptr(std), _stream()),
              "hd_resize_nearest_fwd")


def resize_nearest_bwd(dy, dx, std=None, accumulate=False):
    n, c, hi, wi = dx.shape
    ho, wo = dy.shape[-2:]
    with _Timed("resize_nearest_bwd"):
        check(_lib.load().hd_resize_nearest_bwd(_ptr(dy), _ptr(dx), n, c, hi, wi, ho, wo, _ptr(std), int(accumulate), _stream()),
              "hd_resize_nearest_bwd")


def regulariser(kind, hal, rgb, ir, w_rgb, w_ir, loss, dhal=None, grad_scale=1.0, accumulate=False):
    n, _, h, w = hal.shape
    with _Timed("regulariser"):
        check(_lib.load().hd_regulariser({"mse": 0, "l1": 1}[kind], _ptr(hal), _ptr(rgb), _ptr(ir), w_rgb, w_ir, n, h, w, _ptr(loss),
                                         _ptr(dhal), grad_scale, int(accumulate), _stream()), "hd_regulariser")


NMS_MAX_BOXES = 8192     # per problem (shared-memory budget of the scan kernel)


def nms(boxes, scores, iou_threshold):
    """Drop-in for ``torchvision.ops.nms`` on CUDA fp32 boxes: same stable descending sort, same fp32 IoU predicate, same
    greedy rule (include/hallucidet_b200.h: hd_nms), so the returned indices are identical -- without torchvision's
    one-box-at-a-time mask walk."""
    global LAUNCHES
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    assert boxes.is_cuda and boxes.dtype == torch.float32 and boxes.shape[1] == 4 and n <= NMS_MAX_BOXES
    order = torch.sort(scores, dim=0, descending=True, stable=True)[1]
    sorted_boxes = boxes.index_select(0, order).contiguous()
    mask_ws = torch.empty(n * ((n + 63) // 64), dtype=torch.int64, device=boxes.device)
    keep = torch.empty(n, dtype=torch.bool, device=boxes.device)
    offsets = (ctypes.c_int * 2)(0, n)
    with _Timed("nms"):
        check(_lib.load().hd_nms(_ptr(sorted_boxes), offsets, None, 1, float(iou_threshold), _ptr(mask_ws), _ptr(keep), _stream()), "hd_nms")
    LAUNCHES += 1            # two kernels: pairwise mask + scan
    return order.masked_select(keep)


def nms_sorted_batch(sorted_boxes_list, iou_threshold):
    """Several independent NMS problems (boxes already sorted by descending score) in one pair of launches; returns the
    boolean keep vector of every problem."""
    ns = [int(b.shape[0]) for b in sorted_boxes_list]
    if sum(ns) == 0:
        return [torch.empty(0, dtype=torch.bool, device=b.device) for b in sorted_boxes_list]
    allb = torch.cat([b.reshape(-1, 4) for b in sorted_boxes_list], 0).contiguous()
    offs = [0]
    for n in ns:
        offs.append(offs[-1] + n)
    return list(nms_sorted_flat(allb, offs, iou_threshold).split(ns))


def nms_sorted_flat(sorted_boxes, offsets, iou_threshold, counts=None, valid=None):
    """hd_nms over problems laid back to back in ``sorted_boxes`` [total, 4]; ``offsets`` = host list of problems+1 slot
    offsets; ``counts`` = optional int32 device tensor with the live box count of each problem (no host sync needed);
    ``valid`` = optional bool / uint8 device tensor [total]: boxes with 0 take no part, wherever they sit in their problem."""
    global LAUNCHES
    ns = [offsets[i + 1] - offsets[i] for i in range(len(offsets) - 1)]
    assert sorted_boxes.is_cuda and sorted_boxes.dtype == torch.float32 and sorted_boxes.is_contiguous()
    assert all(0 <= n <= NMS_MAX_BOXES for n in ns) and offsets[-1] == sorted_boxes.shape[0]
    assert counts is None or (counts.dtype == torch.int32 and counts.is_cuda and counts.numel() == len(ns))
    assert valid is None or (valid.dtype in (torch.bool, torch.uint8) and valid.is_cuda and valid.is_contiguous()
                             and valid.numel() == sorted_boxes.shape[0])
    dev = sorted_boxes.device
    mask_ws = torch.empty(max(1, sum(n * ((n + 63) // 64) for n in ns)), dtype=torch.int64, device=dev)
    keep = torch.empty(offsets[-1], dtype=torch.bool, device=dev)
    c_off = (ctypes.c_int * len(offsets))(*offsets)
    with _Timed("nms"):
        check(_lib.load().hd_nms_valid(_ptr(sorted_boxes), c_off, _ptr(counts), _ptr(valid), len(ns), float(iou_threshold), _ptr(mask_ws),
                                       _ptr(keep), _stream()), "hd_nms")
    LAUNCHES += 1
    return keep


class DeviceRng:
    """The default CUDA generator's Philox offset, continued in device memory by hd_sample_balanced so that data-dependent
    draws need no host round trip.  ``begin()`` adopts the host generator's (seed, offset) whenever it differs from what this
    object last saw there (the generator was reseeded or used by someone else); ``sync_host()`` -- called where the step
    syncs with the device anyway -- writes the device-side offset back into the generator."""
    _per_device = {}

    @classmethod
    def get(cls, device):
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        r = cls._per_device.get(idx)
        if r is None:
            r = cls._per_device[idx] = cls(idx)
        return r

    def __init__(self, idx):
        self.idx = idx
        self.dev = torch.device("cuda", idx)
        self.buf = torch.zeros(2, dtype=torch.int64, device=self.dev)      # ping-pong: kernels read one word, write the other
        self.cur = 0
        self.seed = None
        self.host_seen = None
        self.pending = False
        self.ws = {}

    def generator(self):
        if len(torch.cuda.default_generators) <= self.idx:
            torch.cuda.init()
        return torch.cuda.default_generators[self.idx]

    def begin(self):
        g = self.generator()
        state = (int(g.initial_seed()), int(g.get_offset()))
        if state != self.host_seen:
            self.seed = state[0]
            self.buf[self.cur].fill_(state[1])
            self.host_seen = state
            self.pending = False
        return self.seed

    _PENDING_MARK = 1 << 40

    def mark_pending(self):
        """After a device-side draw torch's generator is behind the device.  Its offset is moved to an out-of-the-way marker
        until ``sync_host()``: re-seeding the generator -- even to exactly the state this object adopted last -- is then always
        visible to ``begin()`` as a changed state."""
        if not self.pending:
            g = self.generator()
            marked = int(g.get_offset()) + self._PENDING_MARK
            g.set_offset(marked)
            self.host_seen = (int(g.initial_seed()), marked)
            self.pending = True

    def sync_host(self):
        """Mirror the device-side offset into torch's generator (one 8-byte device->host read; blocks until the stream that
        ran the last draw reaches it)."""
        if not self.pending:
            return
        g = self.generator()
        if (int(g.initial_seed()), int(g.get_offset())) != self.host_seen:      # re-seeded meanwhile: the device chain is obsolete
            self.pending = False
            self.host_seen = None
            return
        off = int(self.buf[self.cur].item())
        g.set_offset(off)
        self.host_seen = (int(g.initial_seed()), off)
        self.pending = False

    def workspace(self, batch):
        w = self.ws.get(batch)
        if w is None:
            n = int(_lib.load().hd_sample_balanced_workspace_bytes(batch))
            if n <= 0:
                raise RuntimeError("hd_sample_balanced_workspace_bytes failed")
            w = self.ws[batch] = torch.empty(n, dtype=torch.uint8, device=self.dev)
        return w


def sample_balanced(labels, batch_size_per_image, positive_fraction):
    """det_utils.BalancedPositiveNegativeSampler for labels [B, N] (float32 or int64; >= 1 positive, 0 negative, else ignored),
    drawn on the device from the default CUDA generator's Philox stream: the same selection torchvision's per-image loop
    makes, without reading the data-dependent randperm sizes on the host.  Returns (sampled [B, N] uint8: 1 = positive drawn,
    2 = negative drawn, 0 = not drawn; counts [B, 4] int32 = positives, negatives, drawn positives, drawn negatives)."""
    global LAUNCHES
    assert labels.is_cuda and labels.dim() == 2 and labels.is_contiguous() and labels.dtype in (torch.float32, torch.int64)
    B, N = labels.shape
    rng = DeviceRng.get(labels.device)
    seed = rng.begin()
    sampled = torch.empty(B, N, dtype=torch.uint8, device=labels.device)
    counts = torch.empty(B, 4, dtype=torch.int32, device=labels.device)
    ws = rng.workspace(B)
    src, dst = rng.buf[rng.cur:rng.cur + 1], rng.buf[1 - rng.cur:2 - rng.cur]
    with _Timed("sample_balanced"):
        check(_lib.load().hd_sample_balanced(_ptr(labels), 0 if labels.dtype == torch.float32 else 1, B, N, int(batch_size_per_image),
                                             int(batch_size_per_image * positive_fraction), seed, _ptr(src), _ptr(dst), _ptr(sampled),
                                             _ptr(counts), _ptr(ws), ws.numel(), _stream()), "hd_sample_balanced")
    rng.cur = 1 - rng.cur
    rng.mark_pending()
    LAUNCHES += 3
    return sampled, counts


def roi_match_labels(props, n_props, gt, gt_present, gt_labels, low_threshold, high_threshold):
    """hd_roi_match_labels: candidates [proposals (padded, live below n_props) | ground truth] of every image -> (labels [B, T + G]
    int64: class / 0 background / -1 ignored or padding, matched [B, T + G] int64).  box_iou + Matcher (no low-quality matches)
    + the label rules of RoIHeads.assign_targets_to_proposals, bit-identical to the PyTorch operator chain."""
    global LAUNCHES
    B, T = props.shape[:2]
    G = gt.shape[1]
    assert props.dtype == gt.dtype == torch.float32 and props.is_contiguous() and gt.is_contiguous() and props.is_cuda
    assert n_props.dtype == torch.int64 and gt_labels.dtype == torch.int64 and gt_present.dtype == torch.bool
    assert gt_present.is_contiguous() and gt_labels.is_contiguous() and 1 <= G <= 64
    labels = torch.empty(B, T + G, dtype=torch.int64, device=props.device)
    matched = torch.empty(B, T + G, dtype=torch.int64, device=props.device)
    with _Timed("roi_match_labels"):
        check(_lib.load().hd_roi_match_labels(_ptr(props), _ptr(n_props), _ptr(gt), _ptr(gt_present), _ptr(gt_labels), B, T, G,
                                              float(low_threshold), float(high_threshold), _ptr(labels), _ptr(matched), _stream()),
              "hd_roi_match_labels")
    LAUNCHES += 1
    return labels, matched


def roi_gather_samples(flat, counts, props, gt, labels, matched, coder_weights, level_mapper):
    """hd_roi_gather_samples: the drawn candidates -> dict(proposals [S, 4], labels [S], matched [S], image_of [S], regression_targets
    [S, 4], rois [S, 5], levels [S], n_drawn (0-dim int64), per_image [B] int64) -- see include/hallucidet_b200.h."""
    global LAUNCHES
    B, T = props.shape[:2]
    G = gt.shape[1]
    S = flat.numel()
    dev = props.device
    a = HdRoiGatherArgs()
    out = {"proposals": torch.empty(S, 4, device=dev), "labels": torch.empty(S, dtype=torch.int64, device=dev),
           "matched": torch.empty(S, dtype=torch.int64, device=dev), "image_of": torch.empty(S, dtype=torch.int64, device=dev),
           "regression_targets": torch.empty(S, 4, device=dev), "rois": torch.empty(S, 5, device=dev),
           "levels": torch.empty(S, dtype=torch.int64, device=dev), "n_drawn": torch.empty((), dtype=torch.int64, device=dev),
           "per_image": torch.empty(B, dtype=torch.int64, device=dev)}
    assert flat.dtype == torch.int64 and counts.dtype == torch.int32 and flat.is_contiguous() and counts.is_contiguous()
    a.flat, a.counts, a.props, a.gt, a.labels, a.matched = (flat.data_ptr(), counts.data_ptr(), props.data_ptr(), gt.data_ptr(),
                                                            labels.data_ptr(), matched.data_ptr())
    a.batch, a.slots, a.n_gt, a.rows = B, T, G, S
    for i, w in enumerate(coder_weights):
        a.weights[i] = float(w)
    lm = level_mapper
    a.canonical_scale, a.canonical_level, a.eps = float(lm.s0), float(lm.lvl0), float(lm.eps)
    a.k_min, a.k_max = float(lm.k_min), float(lm.k_max)
    a.out_props, a.out_labels, a.out_matched, a.out_image = (out["proposals"].data_ptr(), out["labels"].data_ptr(),
                                                             out["matched"].data_ptr(), out["image_of"].data_ptr())
    a.out_targets, a.out_rois, a.out_levels = out["regression_targets"].data_ptr(), out["rois"].data_ptr(), out["levels"].data_ptr()
    a.out_n_drawn, a.out_per_image = out["n_drawn"].data_ptr(), out["per_image"].data_ptr()
    with _Timed("roi_gather_samples"):
        check(_lib.load().hd_roi_gather_samples(ctypes.byref(a), _stream()), "hd_roi_gather_samples")
    LAUNCHES += 1
    return out


def rpn_concat_preds(preds, a, objectness, deltas, backward=False):
    """hd_rpn_concat_preds: channels-last predictor maps [B, H, W, Cp] (a objectness + 4a delta channels) -> objectness [B * sum HWa, 1]
    and deltas [B * sum HWa, 4] in torchvision's order (``backward``: the adjoint, filling ``preds`` from the two flat tensors)."""
    global LAUNCHES
    n = len(preds)
    B = preds[0].shape[0]
    assert all(p.dtype == torch.float32 and p.is_contiguous() and p.dim() == 4 and p.shape[0] == B for p in preds)
    assert objectness.dtype == deltas.dtype == torch.float32 and objectness.is_contiguous() and deltas.is_contiguous()
    total = B * sum(p.shape[1] * p.shape[2] for p in preds) * a
    assert objectness.numel() == total and deltas.numel() == 4 * total
    ptrs = (ctypes.c_void_p * n)(*[p.data_ptr() for p in preds])
    hw = (ctypes.c_int * n)(*[p.shape[1] * p.shape[2] for p in preds])
    cp = (ctypes.c_int * n)(*[p.shape[3] for p in preds])
    with _Timed("rpn_concat_preds"):
        check(_lib.load().hd_rpn_concat_preds(ptrs, hw, cp, n, B, int(a), _ptr(objectness), _ptr(deltas), 1 if backward else 0, _stream()),
              "hd_rpn_concat_preds")
    LAUNCHES += 1


def rpn_assign_targets(anchors, gt, gt_present, low_threshold, high_threshold, allow_low_quality, coder_weights):
    """hd_rpn_assign_targets: anchors [A, 4] (shared by the batch), gt [B, G, 4] padded + presence mask -> (labels [B, A] float32:
    1 / 0 / -1, regression_targets [B * A, 4]); box_iou + Matcher + encode_boxes, bit-identical to the operator chain."""
    global LAUNCHES
    A, (B, G) = anchors.shape[0], gt.shape[:2]
    assert anchors.dtype == gt.dtype == torch.float32 and anchors.is_contiguous() and gt.is_contiguous()
    assert gt_present.dtype == torch.bool and gt_present.is_contiguous() and 1 <= G <= 64
    labels = torch.empty(B, A, device=anchors.device)
    targets = torch.empty(B * A, 4, device=anchors.device)
    ws = torch.empty(B * G, dtype=torch.int32, device=anchors.device)
    w = (ctypes.c_float * 4)(*[float(x) for x in coder_weights])
    with _Timed("rpn_assign_targets"):
        check(_lib.load().hd_rpn_assign_targets(_ptr(anchors), _ptr(gt), _ptr(gt_present), B, A, G, float(low_threshold),
                                                float(high_threshold), int(bool(allow_low_quality)), w, _ptr(ws), _ptr(labels),
                                                _ptr(targets), _stream()), "hd_rpn_assign_targets")
    LAUNCHES += 2
    return labels, targets


def rpn_decode_selected(objectness, deltas, anchors, idx, coder_weights, xform_clip, image_size, min_size, score_thresh):
    """hd_rpn_decode_selected: decode + sigmoid + clip + size / score tests for the candidates ``idx`` [B, M] only.
    objectness [B, A] logits, deltas [B * A, 4], anchors [A, 4] -> (boxes [B, M, 4], scores [B, M], valid [B, M] bool)."""
    global LAUNCHES
    B, M = idx.shape
    A = anchors.shape[0]
    assert objectness.dtype == deltas.dtype == anchors.dtype == torch.float32 and idx.dtype == torch.int64
    assert objectness.is_contiguous() and deltas.is_contiguous() and anchors.is_contiguous() and idx.is_contiguous()
    assert objectness.numel() == B * A and deltas.numel() == 4 * B * A
    boxes = torch.empty(B, M, 4, device=idx.device)
    scores = torch.empty(B, M, device=idx.device)
    valid = torch.empty(B, M, dtype=torch.bool, device=idx.device)
    w = (ctypes.c_float * 4)(*[float(x) for x in coder_weights])
    with _Timed("rpn_decode_selected"):
        check(_lib.load().hd_rpn_decode_selected(_ptr(objectness), _ptr(deltas), _ptr(anchors), _ptr(idx), B, A, M, w, float(xform_clip),
                                                 float(image_size[1]), float(image_size[0]), float(min_size), float(score_thresh),
                                                 _ptr(boxes), _ptr(scores), _ptr(valid), _stream()), "hd_rpn_decode_selected")
    LAUNCHES += 1
    return boxes, scores, valid


def fastrcnn_loss(class_logits, box_regression, labels, regression_targets, beta=1 / 9):
    """hd_fastrcnn_loss: (losses [2] = classification, box regression; d/d class_logits; d/d box_regression) in one launch."""
    global LAUNCHES
    S, C = class_logits.shape
    assert class_logits.dtype == box_regression.dtype == regression_targets.dtype == torch.float32 and labels.dtype == torch.int64
    assert class_logits.is_contiguous() and box_regression.is_contiguous() and regression_targets.is_contiguous() and labels.is_contiguous()
    assert tuple(box_regression.shape) == (S, 4 * C) and tuple(regression_targets.shape) == (S, 4) and labels.numel() == S
    losses = torch.empty(2, device=class_logits.device)
    g_logits, g_box = torch.empty_like(class_logits), torch.empty_like(box_regression)
    with _Timed("fastrcnn_loss"):
        check(_lib.load().hd_fastrcnn_loss(_ptr(class_logits), _ptr(box_regression), _ptr(labels), _ptr(regression_targets), S, C, float(beta),
                                           _ptr(losses), _ptr(g_logits), _ptr(g_box), _stream()), "hd_fastrcnn_loss")
    LAUNCHES += 1
    return losses, g_logits, g_box


def rpn_loss(objectness, pred_bbox_deltas, labels, regression_targets, flat, sampled, counts, beta=1 / 9):
    """hd_rpn_loss: (losses [2] = objectness, box regression; dense d/d objectness; dense d/d pred_bbox_deltas) in one launch
    (+ the two zero fills)."""
    global LAUNCHES
    n = objectness.numel()
    assert objectness.dtype == pred_bbox_deltas.dtype == labels.dtype == regression_targets.dtype == torch.float32
    assert objectness.is_contiguous() and pred_bbox_deltas.is_contiguous() and labels.is_contiguous() and regression_targets.is_contiguous()
    assert pred_bbox_deltas.numel() == 4 * n and labels.numel() == n and regression_targets.numel() == 4 * n
    assert flat.dtype == torch.int64 and sampled.dtype == torch.uint8 and sampled.numel() == n and counts.dtype == torch.int32
    losses = torch.empty(2, device=objectness.device)
    g_obj, g_deltas = torch.zeros_like(objectness), torch.zeros_like(pred_bbox_deltas)
    with _Timed("rpn_loss"):
        check(_lib.load().hd_rpn_loss(_ptr(objectness), _ptr(pred_bbox_deltas), _ptr(labels), _ptr(regression_targets), _ptr(flat),
                                      _ptr(sampled), _ptr(counts), counts.shape[0], flat.numel(), float(beta), _ptr(losses), _ptr(g_obj),
                                      _ptr(g_deltas), _stream()), "hd_rpn_loss")
    LAUNCHES += 1
    return losses, g_obj, g_deltas


def roi_align_bwd(grad_out, rois, input_shape, spatial_scale, sampling_ratio):
    """Gradient of torchvision.ops.roi_align(aligned=False) w.r.t. its fp32 NCHW input of shape ``input_shape``, for grad_out
    [K, C, PH, PW]: channels-last vector reductions (hd_roi_align_bwd_nhwc) + one layout conversion.  Returns NCHW fp32."""
    global LAUNCHES
    k, c, ph, pw = grad_out.shape
    n, c2, h, w = input_shape
    assert c == c2 and grad_out.dtype == torch.float32 and rois.dtype == torch.float32
    assert grad_out.is_contiguous() and rois.is_contiguous() and tuple(rois.shape) == (k, 5)
    scratch = torch.zeros(n, h, w, c, dtype=torch.float32, device=grad_out.device)
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=grad_out.device)
    with _Timed("roi_align_bwd"):
        check(_lib.load().hd_roi_align_bwd_nhwc(_ptr(grad_out), _ptr(rois), _ptr(scratch), k, c, h, w, ph, pw, float(spatial_scale),
                                                int(sampling_ratio), _stream()), "hd_roi_align_bwd_nhwc")
        check(_lib.load().hd_nhwc_to_nchw_f32(_ptr(scratch), _ptr(out), n, c, h, w, _stream()), "hd_nhwc_to_nchw_f32")
    LAUNCHES += 1
    return out


def nchw_to_nhwc_f32(x):
    """fp32 [N, C, H, W] -> fp32 [N, H, W, C] (channels-last copy for the RoIAlign forward)."""
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty(n, h, w, c, dtype=torch.float32, device=x.device)
    with _Timed("nchw_to_nhwc_f32"):
        check(_lib.load().hd_nchw_to_nhwc_f32(_ptr(x), _ptr(y), n, c, h, w, _stream()), "hd_nchw_to_nhwc_f32")
    return y


def roi_align_fwd(feat_nhwc, rois, output_size, spatial_scale, sampling_ratio):
    """torchvision.ops.roi_align(aligned=False) forward on a channels-last fp32 feature map; returns [K, C, PH, PW]."""
    n, h, w, c = feat_nhwc.shape
    k = rois.shape[0]
    assert feat_nhwc.dtype == torch.float32 and feat_nhwc.is_contiguous() and rois.dtype == torch.float32 and rois.is_contiguous()
    out = torch.empty(k, c, int(output_size[0]), int(output_size[1]), dtype=torch.float32, device=rois.device)
    with _Timed("roi_align_fwd"):
        check(_lib.load().hd_roi_align_fwd_nhwc(_ptr(feat_nhwc), _ptr(rois), _ptr(out), k, c, h, w, int(output_size[0]), int(output_size[1]),
                                                float(spatial_scale), int(sampling_ratio), _stream()), "hd_roi_align_fwd_nhwc")
    return out


def _roi_level_table(tensors_nhwc, scales, grads=False):
    table = (_lib.HdRoiLevel * len(tensors_nhwc))()
    for i, (t, sc) in enumerate(zip(tensors_nhwc, scales)):
        assert t.dtype == tensors_nhwc[0].dtype and t.dtype in (torch.float32, torch.bfloat16) and t.is_contiguous()
        if grads:
            table[i].grad_nhwc = _ptr(t)
        else:
            table[i].feat_nhwc = _ptr(t)
        table[i].h, table[i].w, table[i].scale = t.shape[1], t.shape[2], float(sc)
    return table


def roi_align_ml_fwd(feats_nhwc, scales, rois, levels, output_size, sampling_ratio):
    """torchvision MultiScaleRoIAlign's per-level roi_align calls as ONE launch: RoI k is pooled from the channels-last
    level ``levels[k]`` (device int64).  Returns [K, C, PH, PW] fp32, rows in RoI order (bit-identical to the per-level op)."""
    k, c = rois.shape[0], feats_nhwc[0].shape[3]
    assert rois.dtype == torch.float32 and rois.is_contiguous() and levels.dtype == torch.int64 and levels.is_contiguous()
    out = torch.empty(k, c, int(output_size[0]), int(output_size[1]), dtype=torch.float32, device=rois.device)
    table = _roi_level_table(feats_nhwc, scales)
    fn = _lib.load().hd_roi_align_ml_fwd_bf16 if feats_nhwc[0].dtype == torch.bfloat16 else _lib.load().hd_roi_align_ml_fwd
    with _Timed("roi_align_fwd"):
        check(fn(table, len(feats_nhwc), _ptr(rois), _ptr(levels), _ptr(out), k, c, int(output_size[0]),
                 int(output_size[1]), int(sampling_ratio), _stream()), "hd_roi_align_ml_fwd")
    return out


def roi_align_ml_bwd(grad_out, rois, levels, shapes, scales, sampling_ratio, channels_last=None):
    """Gradients of roi_align_ml_fwd w.r.t. the NCHW fp32 feature maps of ``shapes``: one reduction launch into zeroed
    channels-last scratch for all levels, then one layout conversion per level -- none for the levels flagged
    ``channels_last``, whose gradient is returned as the NCHW view of the scratch."""
    global LAUNCHES
    k, c, ph, pw = grad_out.shape
    assert grad_out.dtype == torch.float32 and grad_out.is_contiguous()
    channels_last = channels_last or [False] * len(shapes)
    scratch = [torch.zeros(n, h, w, c2, dtype=torch.float32, device=grad_out.device) for (n, c2, h, w) in shapes]
    outs = [sc.permute(0, 3, 1, 2) if cl else torch.empty(tuple(sh), dtype=torch.float32, device=grad_out.device)
            for sc, sh, cl in zip(scratch, shapes, channels_last)]
    table = _roi_level_table(scratch, scales, grads=True)
    with _Timed("roi_align_bwd"):
        check(_lib.load().hd_roi_align_ml_bwd(table, len(shapes), _ptr(grad_out), _ptr(rois), _ptr(levels), k, c, ph, pw,
                                              int(sampling_ratio), _stream()), "hd_roi_align_ml_bwd")
        for sc, o, (n, c2, h, w), cl in zip(scratch, outs, shapes, channels_last):
            if not cl:
                check(_lib.load().hd_nhwc_to_nchw_f32(_ptr(sc), _ptr(o), n, c2, h, w, _stream()), "hd_nhwc_to_nchw_f32")
    LAUNCHES += 1
    return outs
