"""B200-native drop-in for the frozen detector backbone (torchvision ResNet-50-FPN, ``BackboneWithFPN``).

Keeps the interface the reference's eval-forward functions use (src/utils/eval_forward_fasterrcnn.py:55,
src/utils/eval_forward_retinanet.py:125): ``backbone(Tensor[B,3,S,S] fp32) -> OrderedDict[str, Tensor[B,256,.,.]]``
with keys '0','1','2','3','pool' (Faster R-CNN) or '0','1','2','p6','p7' (RetinaNet), attribute ``out_channels``,
sub-modules ``body`` / ``fpn`` holding the original (frozen) parameters under their torchvision names.

The weights are frozen: BatchNorm is an affine constant folded into the bf16 GEMM operands once
(TV ops/misc.py:54-63), the forward runs conv+bias+ReLU(+residual) fused kernels, and the backward computes
ONLY the input gradient (dgrad) -- never weight gradients -- with the ReLU masks fused into the dgrad epilogues.
No PyTorch/CPU fallback.
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import ops

# FPN outputs / their gradients cross the module boundary as logical NCHW fp32 tensors with CHANNELS-LAST strides: the conv
# epilogue writes them that way (no transpose), cuDNN's RPN / RetinaNet head convolutions and the RoIAlign kernels read
# channels-last natively (no nchwToNhwc / nhwcToNchw passes), and the incoming gradients are a plain cast to bf16.
CHANNELS_LAST_FEATURES = os.environ.get("HD_CL_FEATURES", "1") == "1"
FUSED_STEM = os.environ.get("HD_FUSED_STEM", "1") != "0"       # halo-patch 7x7 stem (csrc/stem_conv.cu) instead of im2col + GEMM
EVAL_CHUNK = int(os.environ.get("HD_EVAL_BACKBONE_CHUNK", "16"))      # images per engine pass when no gradient is needed


class _FConv:
    """A frozen conv (+folded BN): packed operands, folded bias, output buffer."""

    def __init__(self, eng, conv, bn, in_hw, relu, stem=False):
        self.conv, self.relu = conv, relu
        self.cout, self.cin = conv.out_channels, conv.in_channels
        self.k, self.stride = conv.kernel_size[0], conv.stride[0]
        self.h_in, self.w_in = in_hw
        # stride-2 convs on odd feature maps (S=300: 75, 19) run on an even zero-padded copy: out = ceil(in / 2) as torch
        self.pad_in = self.stride == 2 and (in_hw[0] % 2 == 1 or in_hw[1] % 2 == 1) and not stem
        self.hp, self.wp = in_hw[0] + in_hw[0] % 2, in_hw[1] + in_hw[1] % 2
        self.h, self.w = (in_hw[0] + self.stride - 1) // self.stride, (in_hw[1] + self.stride - 1) // self.stride
        dev = eng.device
        w = conv.weight.detach().float().contiguous()
        if bn is not None:
            eps = bn.eps
            scale = (bn.weight.detach() * torch.rsqrt(bn.running_var + eps)).float().contiguous()
            self.bias = (bn.bias.detach() - bn.running_mean * scale).float().contiguous()
        else:
            scale = None
            self.bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
        self.scale = scale
        if stem:
            self.packed = ops.PackedConv(self.cout, 3, 7, dev, need_dgrad=False, need_t=True, k_pad=ops.STEM_KPAD).pack(w, scale)
        else:
            self.packed = ops.PackedConv(self.cout, self.cin, self.k, dev, need_dgrad=True).pack(w, scale)
        self.y = eng.new_act(self.h, self.w, self.cout)
        if self.pad_in:
            self.x_pad = eng.new_act(self.hp, self.wp, self.cin)
            self.dx_pad = None


class FrozenBackbone(nn.Module):
    def __init__(self, body, fpn, variant):
        super().__init__()
        self.body, self.fpn = body, fpn
        self.variant = variant
        self.out_channels = fpn.layer_blocks[0][0].out_channels if isinstance(fpn.layer_blocks[0], nn.Sequential) else 256
        for p in self.parameters():
            p.requires_grad_(False)
        self._engines = {}
        self._last_engine = None
        self.use_cuda_graph = False

    @classmethod
    def from_torchvision(cls, backbone):
        """backbone: torchvision ``BackboneWithFPN`` (``detector.backbone``) with its weights already loaded."""
        from torchvision.ops.feature_pyramid_network import LastLevelMaxPool, LastLevelP6P7
        extra = backbone.fpn.extra_blocks
        if isinstance(extra, LastLevelMaxPool):
            variant = "fasterrcnn"
        elif isinstance(extra, LastLevelP6P7):
            if not extra.use_P5:
                raise NotImplementedError("LastLevelP6P7 on C5 (retinanet v2) is not implemented")
            variant = "retinanet"
        else:
            raise NotImplementedError(f"FPN extra block {type(extra).__name__}")
        return cls(backbone.body, backbone.fpn, variant)

    def bf16_features(self):
        """The bf16 NHWC pyramid of the LATEST forward, in the order of the returned dict (valid until the next forward of the
        same engine): what the frozen-head kernels read instead of the fp32 copies."""
        last = getattr(self, "_last_engine", None)
        if last is None or last[0].generation != last[1]:
            return None
        eng = last[0]
        feats = [lv["layer"].y for lv in eng.levels] + [e["conv"].y for e in eng.extra]
        if self.variant == "fasterrcnn":
            feats.append(eng.pool_bf16)
        return feats

    def refold(self):
        """Re-fold BatchNorm / re-pack operands after the frozen weights were (re)loaded."""
        self._engines.clear()

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._engines.clear()

    def _engine(self, x):
        # inputs that carry no gradient (the reference's extra RGB / IR detector passes, train_hallucidet.py:183,186)
        # get their own activation set so that they do not overwrite what the pending backward needs
        needs_grad = bool(x.requires_grad and torch.is_grad_enabled())
        key = (tuple(x.shape), x.device.index, self.body.conv1.weight.data_ptr(), needs_grad)
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= 4:
                self._engines.clear()
            eng = _BackboneEngine(self, x.shape[0], x.shape[2], x.shape[3], x.device)
            self._engines[key] = eng
        return eng

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("hallucidet_b200.FrozenBackbone runs only on a CUDA (B200) device; there is no CPU path")
        if not (x.requires_grad and torch.is_grad_enabled()) and x.shape[0] > EVAL_CHUNK:
            # gradient-free pass over a large batch (inference): chunks through one engine, concatenated per level
            parts = [self.forward(c) for c in x.split(EVAL_CHUNK)]
            self._last_engine = None                       # (the bf16 pyramid only holds the last chunk)
            return OrderedDict((k, torch.cat([p[k] for p in parts], 0)) for k in parts[0])
        x = x.contiguous().float()
        eng = self._engine(x)
        self._last_engine = (eng, eng.generation + 1)
        outs = _BackboneFunction.apply(x, eng)
        names = list(eng.level_names)
        res = OrderedDict(zip(names, outs))
        if self.variant == "fasterrcnn":
            # LastLevelMaxPool: max_pool2d(k=1, s=2) == stride-2 subsample (TV ops/feature_pyramid_network.py:207-221)
            res["pool"] = res[names[-1]][:, :, ::2, ::2]
        return res


class _BackboneFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng):
        ctx.eng = eng
        outs = eng.forward(x, clone=not getattr(eng.m, "static_outputs", False))
        ctx.generation = eng.generation
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        eng = ctx.eng
        if ctx.generation != eng.generation:
            raise RuntimeError("hallucidet_b200.FrozenBackbone: activations of this forward were overwritten by a later forward")
        return eng.backward(grads), None


class _BackboneEngine:
    def __init__(self, module, B, H, W, device):
        if H % 4 != 0 or W % 4 != 0 or H < 64 or W < 64:
            raise NotImplementedError(f"FrozenBackbone input {H}x{W}: height and width must be multiples of 4 and >= 64 "
                                      "(stem conv and max-pool run on even maps); HalluciDet uses 300 and 640")
        self.m, self.B, self.H, self.W, self.device = module, B, H, W, device
        self.variant = module.variant
        body, fpn = module.body, module.fpn
        self.generation = 0
        self.graphs = {}
        self.x_in = torch.empty(B, 3, H, W, device=device)
        h2, w2 = H // 2, W // 2
        # stem
        self.stem = _FConv.__new__(_FConv)
        st = self.stem
        _FConv.__init__(st, self, body.conv1, body.bn1, (H, W), relu=True, stem=True)
        self.patches = torch.empty(1, 1, B * h2 * w2, ops.STEM_KPAD, dtype=torch.bfloat16, device=device)
        self.p0 = self.new_act(h2 // 2, w2 // 2, 64)
        self.p0_idx = torch.empty(B, h2 // 2, w2 // 2, 64, dtype=torch.uint8, device=device)   # max-pool arg-max (ReLU mask folded in)
        # body
        self.blocks = []
        hw = (h2 // 2, w2 // 2)
        self.c_out = {}
        for li in range(1, 5):
            layer = getattr(body, f"layer{li}")
            for b, blk in enumerate(layer):
                c1 = _FConv(self, blk.conv1, blk.bn1, hw, relu=True)
                c2 = _FConv(self, blk.conv2, blk.bn2, hw, relu=True)
                ohw = (c2.h, c2.w)
                c3 = _FConv(self, blk.conv3, blk.bn3, ohw, relu=True)
                cd = _FConv(self, blk.downsample[0], blk.downsample[1], hw, relu=False) if blk.downsample is not None else None
                self.blocks.append(dict(c1=c1, c2=c2, c3=c3, cd=cd, li=li, first=(b == 0)))
                hw = ohw
            self.c_out[li] = self.blocks[-1]["c3"].y
        # FPN
        ret = [1, 2, 3, 4] if self.variant == "fasterrcnn" else [2, 3, 4]
        self.levels = []
        for i, li in enumerate(ret):
            c = self.c_out[li]
            inner = _FConv(self, fpn.inner_blocks[i][0], None, (c.shape[1], c.shape[2]), relu=False)
            layer = _FConv(self, fpn.layer_blocks[i][0], None, (c.shape[1], c.shape[2]), relu=False)
            out = self._new_out(layer.cout, layer.h, layer.w)
            self.levels.append(dict(li=li, inner=inner, layer=layer, out=out, name=str(i)))
        self.level_names = [lv["name"] for lv in self.levels]
        self.extra = []
        if self.variant == "retinanet":
            top = self.levels[-1]["layer"]
            p6 = _FConv(self, fpn.extra_blocks.p6, None, (top.h, top.w), relu=False)
            self.p6_relu = self.new_act(p6.h, p6.w, p6.cout)
            p7 = _FConv(self, fpn.extra_blocks.p7, None, (p6.h, p6.w), relu=False)
            self.extra = [dict(conv=p6, out=self._new_out(256, p6.h, p6.w), name="p6"),
                          dict(conv=p7, out=self._new_out(256, p7.h, p7.w), name="p7")]
            self.level_names += ["p6", "p7"]
            self.ones = torch.ones(256, device=device)
            self.zeros = torch.zeros(256, device=device)
        else:
            top = self.levels[-1]["layer"]
            self.pool_bf16 = self.new_act((top.h + 1) // 2, (top.w + 1) // 2, top.cout)
        self.grad_bufs = {}
        self.dx = torch.empty(B, 3, H, W, device=device)
        self.dP_in = [torch.empty_like(lv["out"]) for lv in self.levels] + [torch.empty_like(e["out"]) for e in self.extra]

    def _new_out(self, c, h, w):
        """fp32 [B, C, H, W] output buffer (channels-last strides when CHANNELS_LAST_FEATURES)."""
        self.cl = CHANNELS_LAST_FEATURES and c % 16 == 0
        if self.cl:
            return torch.empty(self.B, h, w, c, device=self.device).permute(0, 3, 1, 2)
        return torch.empty(self.B, c, h, w, device=self.device)

    def _grad_to_nhwc_bf16(self, src, dst, accumulate=False):
        """fp32 gradient of an output (layout of the output buffers) -> bf16 NHWC ``dst`` (captured in the CUDA graph)."""
        if self.cl:
            nhwc = src.permute(0, 2, 3, 1)
            if accumulate:
                dst.add_(nhwc)
            else:
                dst.copy_(nhwc)
        else:
            ops.nchw_f32_to_nhwc_bf16(src, dst, accumulate=accumulate)

    def new_act(self, h, w, c):
        return torch.empty(self.B, h, w, c, dtype=torch.bfloat16, device=self.device)

    def gbuf(self, key, like):
        t = self.grad_bufs.get(key)
        if t is None:
            t = torch.empty_like(like)
            self.grad_bufs[key] = t
        return t

    def _run(self, kind, fn):
        if not self.m.use_cuda_graph:
            fn()
            return
        state = self.graphs.get(kind)
        if state is None:
            fn()
            self.graphs[kind] = "warm"
        elif state == "warm":
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self.graphs[kind] = g
            g.replay()
        else:
            state.replay()

    # ---- forward -------------------------------------------------------------------------------------
    def _conv(self, c, x, add=None, out_f32=None, store_bf16=True):
        if c.pad_in:
            ops.pad_hw(x, c.x_pad)
            x = c.x_pad
        ops.conv_fwd(ops.conv_args(x, c.y, c.packed.w_fwd, k=c.k, stride=c.stride, bias=c.bias, add=add, relu=c.relu,
                                   out_f32=out_f32, out_f32_channels=c.cout if out_f32 is not None else 0, store_bf16=store_bf16,
                                   out_f32_nhwc=out_f32 is not None and self.cl))

    def forward(self, x, clone=True):
        self.x_in.copy_(x)
        self._run("fwd", self._forward_impl)
        self.generation += 1
        if not clone:      # the engine's own buffers (valid until its next forward): FrozenBackbone.static_outputs
            return [lv["out"] for lv in self.levels] + [e["out"] for e in self.extra]
        return [lv["out"].clone() for lv in self.levels] + [e["out"].clone() for e in self.extra]

    def _forward_impl(self):
        st = self.stem
        if FUSED_STEM:        # halo-patch stem kernel straight from the fp32 master weight (frozen BatchNorm scale folded in the kernel)
            ops.stem_fwd(self.x_in, st.conv.weight.detach(), st.y, w_scale=st.scale, bias=st.bias, relu=True)
        else:
            ops.stem_im2col(self.x_in, self.patches)
            ops.conv_fwd(ops.conv_args(self.patches, st.y.view(1, 1, -1, 64), st.packed.w_fwd, k=1, bias=st.bias, relu=True, algo_cin=147))
        ops.maxpool_fwd(st.y, self.p0, idx=self.p0_idx, mask_nonpositive=True)
        x = self.p0
        for blk in self.blocks:
            c1, c2, c3, cd = blk["c1"], blk["c2"], blk["c3"], blk["cd"]
            blk["x_in"] = x
            self._conv(c1, x)
            self._conv(c2, c1.y)
            if cd is not None:
                self._conv(cd, x)
                self._conv(c3, c2.y, add=cd.y)
            else:
                self._conv(c3, c2.y, add=x)
            x = c3.y
        # FPN top-down (TV ops/feature_pyramid_network.py:187-197)
        top = len(self.levels) - 1
        for i in range(top, -1, -1):
            lv = self.levels[i]
            self._conv(lv["inner"], self.c_out[lv["li"]])
            if i < top:
                ops.add_nearest_fwd(self.levels[i + 1]["inner"].y, lv["inner"].y)
            # bf16 copy of every pyramid level: the frozen heads (hallucidet_b200/heads.py) read it directly
            self._conv(lv["layer"], lv["inner"].y, out_f32=lv["out"], store_bf16=True)
        if self.extra:
            p6, p7 = self.extra[0]["conv"], self.extra[1]["conv"]
            self._conv(p6, self.levels[top]["layer"].y, out_f32=self.extra[0]["out"])
            ops.bn_apply(p6.y, self.ones, self.zeros, self.p6_relu, relu=True)
            self._conv(p7, self.p6_relu, out_f32=self.extra[1]["out"], store_bf16=True)
        else:
            self.pool_bf16.copy_(self.levels[top]["layer"].y[:, ::2, ::2, :])            # LastLevelMaxPool of the bf16 pyramid

    # ---- backward (input gradient only) ---------------------------------------------------------------
    def backward(self, grads):
        nl = len(self.levels)
        for i, (buf, g) in enumerate(zip(self.dP_in, grads)):
            if self.cl and i < nl:
                # channels-last: the fp32 gradient is cast straight into the graph's static bf16 NHWC buffer (any strides)
                dst = self.gbuf(("gP", i), self.levels[i]["layer"].y)
                if g is None:
                    dst.zero_()
                else:
                    dst.copy_(g.permute(0, 2, 3, 1))
            elif g is None:
                buf.zero_()
            else:
                buf.copy_(g)
        self._run("bwd", self._backward_impl)
        return self.dx.clone()

    def _dgrad(self, c, dy, dx, add=None, mask=None):
        if c.pad_in:
            if c.dx_pad is None:
                c.dx_pad = torch.empty_like(c.x_pad)
            if c.k == 1:
                c.dx_pad.zero_()                         # a 1x1 stride-2 dgrad only writes the (even, even) phase
            ops.conv_dgrad(ops.conv_args(dy, c.dx_pad, c.packed.w_dgrad, k=c.k, stride=c.stride))
            ops.crop_add_mask(c.dx_pad, dx, add=add, mask=mask)
            return
        ops.conv_dgrad(ops.conv_args(dy, dx, c.packed.w_dgrad, k=c.k, stride=c.stride, add=add, mask=mask))

    def _backward_impl(self):
        nl = len(self.levels)
        top = nl - 1
        # gradients of the FPN outputs arrive as fp32 NCHW -> bf16 NHWC
        gP = []
        for i, lv in enumerate(self.levels):
            g = self.gbuf(("gP", i), lv["layer"].y)
            if not self.cl:                                          # (channels-last: filled by backward() before the graph runs)
                ops.nchw_f32_to_nhwc_bf16(self.dP_in[i], g)
            gP.append(g)
        if self.extra:
            p6, p7 = self.extra[0]["conv"], self.extra[1]["conv"]
            g7 = self.gbuf(("gP", "p7"), p7.y)
            self._grad_to_nhwc_bf16(self.dP_in[nl + 1], g7)
            g6 = self.gbuf(("gP", "p6"), p6.y)
            self._dgrad(p7, g7, g6, mask=p6.y)                       # through ReLU(p6)
            self._grad_to_nhwc_bf16(self.dP_in[nl], g6, accumulate=True)
            g5 = self.gbuf(("g5x", 0), gP[top])
            self._dgrad(p6, g6, g5, add=gP[top])
            gP[top] = g5
        # FPN: P_l = layer_l(inner_l);  inner_l = lateral_l(C_l) + nearest(inner_{l+1})
        g_inner = []
        for i in range(nl):
            lv = self.levels[i]
            gi = self.gbuf(("ginner", i), lv["inner"].y)
            self._dgrad(lv["layer"], gP[i], gi)
            if i > 0:
                ops.add_nearest_bwd(g_inner[i - 1], gi, accumulate=True)
            g_inner.append(gi)
        lateral_of = {lv["li"]: (lv["inner"], g_inner[i]) for i, lv in enumerate(self.levels)}
        # body, deepest block first.  g_pre = gradient w.r.t. the pre-ReLU sum of a block (already masked).
        g_next = None                                                  # gradient w.r.t. the current block's OUTPUT from later blocks
        for bi in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[bi]
            c1, c2, c3, cd = blk["c1"], blk["c2"], blk["c3"], blk["cd"]
            out = c3.y
            last_of_layer = bi == len(self.blocks) - 1 or self.blocks[bi + 1]["li"] != blk["li"]
            if last_of_layer and blk["li"] in lateral_of:
                inner, gi = lateral_of[blk["li"]]
                g_pre = self.gbuf(("gpre", bi), out)
                self._dgrad(inner, gi, g_pre, add=g_next, mask=out)
            else:
                g_pre = g_next                                         # already masked by the consumer's epilogue
                if g_pre is None:
                    raise RuntimeError("backbone block without gradient")
            g_h2 = self.gbuf(("gh2", bi), c2.y)
            self._dgrad(c3, g_pre, g_h2, mask=c2.y)
            g_h1 = self.gbuf(("gh1", bi), c1.y)
            self._dgrad(c2, g_h2, g_h1, mask=c1.y)
            x_in = blk["x_in"]
            g_x = self.gbuf(("gx", bi), x_in)
            first_block = bi == 0                                      # its input is the max-pool output: no ReLU mask
            if cd is None:
                self._dgrad(c1, g_h1, g_x, add=g_pre, mask=None if first_block else x_in)
            else:
                self._dgrad(c1, g_h1, g_x)
                self._dgrad(cd, g_pre, g_x, add=g_x, mask=None if first_block else x_in)
            # g_x is masked by (x_in > 0) == pre-ReLU gradient of the previous block -- unless the previous block
            # also feeds an FPN lateral, in which case the lateral dgrad adds its part and masks (handled above
            # through add=g_next with mask=out); there g_next must be the UNMASKED sum, and masking twice is idempotent.
            g_next = g_x
        # stem: max-pool backward with the stem ReLU mask, patch GEMM with W^T, col2im
        st = self.stem
        g_stem = self.gbuf(("g", "stem"), st.y)
        ops.maxpool_bwd(st.y, self.p0, g_next, g_stem, idx=self.p0_idx)
        dpatch = self.gbuf(("dpatch", 0), self.patches)
        ops.conv_fwd(ops.conv_args(g_stem.view(1, 1, -1, 64), dpatch, st.packed.w_t, k=1, algo_cout=147))
        ops.stem_col2im(dpatch, self.dx)
