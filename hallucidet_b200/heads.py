"""Frozen detector heads on the B200 conv kernels.

The RPN head (TV models/detection/rpn.py:15-78, called at src/utils/eval_forward_fasterrcnn.py:76) and the RetinaNet head
(TV models/detection/retinanet.py:87-260, called at src/utils/eval_forward_retinanet.py:131) are towers of
3x3 256->256 convolutions + ReLU shared over the pyramid levels, followed by one predictor convolution.  In HalluciDet the
detector is frozen, so -- exactly like the backbone -- they need the forward and the INPUT gradient only.  Here they run on
``hd_conv_fwd`` / ``hd_conv_dgrad`` (tcgen05 implicit GEMM, bf16 NHWC activations, fp32 accumulation):

  * input = the bf16 NHWC pyramid the backbone kernels already hold (no fp32 -> bf16 round trip),
  * bias + ReLU in the conv epilogue, the ReLU mask in the dgrad epilogue,
  * the RPN's two 1x1 predictors (objectness | box deltas) as ONE 1x1 convolution over 16 padded output channels,
  * predictor outputs written as fp32 channels-last straight from the epilogue (what ``concat_box_prediction_layers`` /
    the RetinaNet permutes flatten without a copy),
  * no weight gradients.

Modules and ``state_dict`` keys are untouched (the engines read the frozen parameters once and re-pack when they move).
There is no PyTorch fallback inside: callers use these only on CUDA tensors coming from ``FrozenBackbone``.
"""
import torch

from . import ops


def _round16(c):
    return (c + 15) // 16 * 16


class _Tower:
    """tower = [(conv3x3 + bias + ReLU)] * n followed by a predictor conv (k = 1 or 3, no activation); weights shared by all levels."""

    def __init__(self, tower_convs, pred_weight, pred_bias, device):
        self.device = device
        self.convs = []
        for conv in tower_convs:
            w = conv.weight.detach().float().contiguous()
            assert tuple(w.shape[2:]) == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.groups == 1
            b = conv.bias.detach().float().contiguous() if conv.bias is not None else torch.zeros(w.shape[0], device=device)
            self.convs.append((ops.PackedConv(w.shape[0], w.shape[1], 3, device).pack(w), b, w.shape[0]))
        pw = pred_weight.detach().float().contiguous()
        self.pred_k = pw.shape[2]
        self.pred_c = pw.shape[0]
        self.pred_cp = _round16(self.pred_c)
        wpad = torch.zeros(self.pred_cp, pw.shape[1], self.pred_k, self.pred_k, device=device)
        wpad[:self.pred_c] = pw
        self.pred = ops.PackedConv(self.pred_cp, pw.shape[1], self.pred_k, device).pack(wpad)
        self.pred_bias = torch.zeros(self.pred_cp, device=device)
        self.pred_bias[:self.pred_c] = pred_bias.detach().float()
        self.scratch = {}
        self.programs = {}

    def _scratch(self, shape):
        """Scratch of one pyramid level (keyed by its [B, H, W, C] shape): nothing here has to survive until the backward."""
        st = self.scratch.get(shape)
        if st is None:
            b, h, w, c = shape
            dev = self.device
            st = self.scratch[shape] = {
                "pred_bf16": torch.empty(b, h, w, self.pred_cp, dtype=torch.bfloat16, device=dev),     # (never written: store_bf16 = 0)
                "dpred": torch.empty(b, h, w, self.pred_cp, dtype=torch.bfloat16, device=dev),
                "g": [torch.empty(b, h, w, co, dtype=torch.bfloat16, device=dev) for _, _, co in self.convs],
                "dx_dummy": torch.empty(b, h, w, c, dtype=torch.bfloat16, device=dev),
            }
        return st

    def forward(self, x, hidden=None, pred=None):
        """x: bf16 NHWC [B, H, W, C] -> (fp32 [B, H, W, pred_cp] channels-last predictor output, hidden activations).
        ``hidden`` / ``pred``: optional pre-allocated outputs (CUDA-graph programs)."""
        st = self._scratch(tuple(x.shape))
        b, hh, ww, _ = x.shape
        h = x
        if hidden is None:                                    # kept for the ReLU mask of the backward
            hidden = [torch.empty(b, hh, ww, co, dtype=torch.bfloat16, device=self.device) for _, _, co in self.convs]
        for (pk, bias, co), out in zip(self.convs, hidden):
            ops.conv_fwd(ops.conv_args(h, out, pk.w_fwd, k=3, bias=bias, relu=True))
            h = out
        if pred is None:
            pred = torch.empty(b, hh, ww, self.pred_cp, device=self.device)
        ops.conv_fwd(ops.conv_args(h, st["pred_bf16"], self.pred.w_fwd, k=self.pred_k, bias=self.pred_bias,
                                   out_f32=pred, out_f32_channels=self.pred_cp, out_f32_nhwc=True, store_bf16=False))
        return pred, hidden

    def backward(self, x_shape, hidden, dpred, dx=None):
        """dpred: fp32 [B, H, W, pred_cp] (any strides) -> fp32 [B, H, W, C] gradient of the level input."""
        st = self._scratch(tuple(x_shape))
        st["dpred"].copy_(dpred)                                            # fp32 -> bf16 NHWC
        g = st["dpred"]
        pk, k = self.pred, self.pred_k
        for i in range(len(self.convs) - 1, -1, -1):
            # gradient w.r.t. the ReLU output of tower conv i, masked by that ReLU in the epilogue
            ops.conv_dgrad(ops.conv_args(g, st["g"][i], pk.w_dgrad, k=k, mask=hidden[i]))
            g = st["g"][i]
            pk, k = self.convs[i][0], 3
        b, hh, ww, c = x_shape
        if dx is None:
            dx = torch.empty(b, hh, ww, c, device=self.device)
        ops.conv_dgrad(ops.conv_args(g, st["dx_dummy"], pk.w_dgrad, k=k, out_f32=dx, out_f32_channels=c, out_f32_nhwc=True,
                                     store_bf16=False))
        return dx


_GRAD_ENABLED_AT_CALL = [False]  # torch.is_grad_enabled() where _TowerFunction.apply was called (see _tower_apply)
USE_CUDA_GRAPH = False          # set by the caller (detection.py: when the backbone runs its kernel programs as CUDA graphs)


class _TowerProgram:
    """The tower over all pyramid levels as two CUDA graphs (forward, input gradient) on static buffers: the levels' launches
    (2-10 per level and direction, each with host-side tensor-map encoding) become one graph launch each -- the detection tail
    is host-bound, so this is time on the critical path.  Keyed by the input pointers (the backbone's static bf16 pyramid)."""

    def __init__(self, tower, feats_bf16):
        self.tower, self.x = tower, list(feats_bf16)
        dev = tower.device
        self.hidden = [[torch.empty(x.shape[0], x.shape[1], x.shape[2], co, dtype=torch.bfloat16, device=dev) for _, _, co in tower.convs]
                       for x in self.x]
        self.pred = [torch.empty(x.shape[0], x.shape[1], x.shape[2], tower.pred_cp, device=dev) for x in self.x]
        self.dpred = [torch.zeros_like(p) for p in self.pred]
        self.dx = [torch.empty(tuple(x.shape), device=dev) for x in self.x]
        self.graphs = {}
        self.generation = 0

    def _run(self, kind, fn):
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

    def forward(self):
        self._run("fwd", lambda: [self.tower.forward(x, h, p) for x, h, p in zip(self.x, self.hidden, self.pred)])
        self.generation += 1
        return tuple(self.pred)

    def backward(self, dpreds, need):
        for buf, dp in zip(self.dpred, dpreds):
            if dp is None:
                buf.zero_()
            elif dp.data_ptr() != buf.data_ptr():
                buf.copy_(dp)
        self._run("bwd", lambda: [self.tower.backward(tuple(x.shape), h, dp, dx)
                                  for x, h, dp, dx in zip(self.x, self.hidden, self.dpred, self.dx)])
        return [dx.permute(0, 3, 1, 2) if n else None for dx, n in zip(self.dx, need)]


class _TowerFunction(torch.autograd.Function):
    """One tower over all pyramid levels.  ``feats``: the fp32 feature maps autograd knows (their gradient is returned);
    ``feats_bf16``: the same values as the backbone's bf16 NHWC buffers, which the kernels read."""

    @staticmethod
    def forward(ctx, tower, feats_bf16, *feats):
        ctx.tower = tower
        ctx.shapes = [tuple(x.shape) for x in feats_bf16]
        ctx.need = [f.requires_grad for f in feats]
        ctx.program = None
        # (grad mode is always off INSIDE Function.forward: the caller's grad mode travels in _GRAD_ENABLED_AT_CALL)
        if USE_CUDA_GRAPH and any(ctx.need) and _GRAD_ENABLED_AT_CALL[0]:
            key = tuple(x.data_ptr() for x in feats_bf16)
            prog = tower.programs.get(key)
            if prog is None:
                if len(tower.programs) >= 4:
                    tower.programs.clear()
                prog = tower.programs[key] = _TowerProgram(tower, feats_bf16)
            ctx.program = prog
            outs = prog.forward()
            ctx.generation = prog.generation
            return outs
        outs = [tower.forward(x) for x in feats_bf16]
        ctx.hidden = [h for _, h in outs]
        return tuple(p for p, _ in outs)

    @staticmethod
    def backward(ctx, *dpreds):
        if ctx.program is not None:
            if ctx.generation != ctx.program.generation:
                raise RuntimeError("hallucidet_b200.heads: the activations of this forward were overwritten by a later forward")
            return (None, None) + tuple(ctx.program.backward(dpreds, ctx.need))
        grads = []
        for shape, need, hidden, dp in zip(ctx.shapes, ctx.need, ctx.hidden, dpreds):
            if not need or dp is None:
                grads.append(None)
                continue
            grads.append(ctx.tower.backward(shape, hidden, dp).permute(0, 3, 1, 2))      # logical NCHW, channels-last strides
        return (None, None) + tuple(grads)


def _tower_apply(tower, feats_bf16, *feats):
    _GRAD_ENABLED_AT_CALL[0] = torch.is_grad_enabled()
    return _TowerFunction.apply(tower, feats_bf16, *feats)


def _frozen(*params):
    return not any(p is not None and p.requires_grad for p in params)


def _tower_key(convs, extra):
    return tuple(c.weight.data_ptr() for c in convs) + tuple(e.data_ptr() for e in extra)


def _conv_of(block):
    """The nn.Conv2d of a torchvision Conv2dNormActivation(conv, ReLU) block (norm_layer=None), or None."""
    if isinstance(block, torch.nn.Sequential) and len(block) == 2 and isinstance(block[0], torch.nn.Conv2d) \
            and isinstance(block[1], torch.nn.ReLU):
        return block[0]
    return None


def rpn_head_tower(head):
    """The (cached) B200 tower of a frozen torchvision ``RPNHead``, or None if its structure is not the stock one."""
    convs = [_conv_of(b) for b in head.conv] if isinstance(head.conv, torch.nn.Sequential) else [None]
    cls, reg = head.cls_logits, head.bbox_pred
    if any(c is None for c in convs) or not (isinstance(cls, torch.nn.Conv2d) and isinstance(reg, torch.nn.Conv2d)):
        return None
    if cls.kernel_size != (1, 1) or reg.kernel_size != (1, 1) or cls.bias is None or reg.bias is None:
        return None
    if not _frozen(cls.weight, cls.bias, reg.weight, reg.bias, *[p for c in convs for p in (c.weight, c.bias)]):
        return None
    key = _tower_key(convs, (cls.weight, reg.weight, cls.bias, reg.bias))
    cached = getattr(head, "_hd_tower", None)
    if cached is None or cached[0] != key:
        tower = _Tower(convs, torch.cat([cls.weight, reg.weight], 0), torch.cat([cls.bias, reg.bias], 0), cls.weight.device)
        cached = head._hd_tower = (key, tower)
    return cached[1]


def rpn_head_forward(head, features, features_bf16, return_static=False):
    """``RPNHead.forward`` (TV rpn.py:61-68) on the B200 kernels -> (objectness list, bbox-delta list), each [B, A | 4A, H, W]
    fp32 views of one channels-last predictor output per level.  ``return_static``: also tell whether the outputs live in the
    static buffers of a CUDA-graph program (same addresses every step: downstream static-shape work can be graphed too)."""
    tower = rpn_head_tower(head)
    a = head.cls_logits.out_channels
    feats_bf16 = list(features_bf16)
    preds = _tower_apply(tower, feats_bf16, *features)
    logits, bbox = [], []
    for p in preds:
        nchw = p.permute(0, 3, 1, 2)
        logits.append(nchw[:, :a])
        bbox.append(nchw[:, a:a + 4 * a])
    if return_static:
        prog = tower.programs.get(tuple(x.data_ptr() for x in feats_bf16))
        return logits, bbox, bool(prog is not None and preds[0].data_ptr() == prog.pred[0].data_ptr())
    return logits, bbox


class _ConcatRPNPreds(torch.autograd.Function):
    """torchvision ``concat_box_prediction_layers`` on the tower's channels-last predictor maps, one launch each way
    (ops.rpn_concat_preds) instead of ~12 copies forward and autograd's ~30 slice / cat / permute kernels backward."""

    @staticmethod
    def forward(ctx, a, out, prog, *preds):
        B = preds[0].shape[0]
        total = B * sum(p.shape[1] * p.shape[2] for p in preds) * a
        if out is None:
            out = (torch.empty(total, 1, device=preds[0].device), torch.empty(total, 4, device=preds[0].device))
        ops.rpn_concat_preds([p.detach() for p in preds], a, out[0], out[1])
        ctx.a, ctx.shapes, ctx.prog = a, [tuple(p.shape) for p in preds], prog
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_obj, g_deltas):
        # straight into the tower program's static gradient buffers when there is one (its backward then skips its copies)
        grads = ctx.prog.dpred if ctx.prog is not None else [torch.empty(s, device=g_obj.device, dtype=torch.float32) for s in ctx.shapes]
        ops.rpn_concat_preds(grads, ctx.a, g_obj.contiguous(), g_deltas.contiguous(), backward=True)
        return (None, None, None) + tuple(grads)


def rpn_head_forward_flat(head, features, features_bf16):
    """``RPNHead.forward`` + ``concat_box_prediction_layers``: (objectness [B * A, 1], deltas [B * A, 4], anchors per level,
    static) -- ``static``: both tensors live at fixed addresses (the tower runs as a CUDA-graph program), so downstream
    fixed-shape work can be captured too."""
    tower = rpn_head_tower(head)
    a = head.cls_logits.out_channels
    feats_bf16 = list(features_bf16)
    preds = _tower_apply(tower, feats_bf16, *features)
    prog = tower.programs.get(tuple(x.data_ptr() for x in feats_bf16))
    static = bool(prog is not None and preds[0].data_ptr() == prog.pred[0].data_ptr())
    out = None
    if static:
        out = getattr(prog, "flat_out", None)
        if out is None:
            total = preds[0].shape[0] * sum(p.shape[1] * p.shape[2] for p in preds) * a
            out = prog.flat_out = (torch.empty(total, 1, device=preds[0].device), torch.empty(total, 4, device=preds[0].device))
    objectness, deltas = _ConcatRPNPreds.apply(a, out, prog if static else None, *preds)
    return objectness, deltas, [p.shape[1] * p.shape[2] * a for p in preds], static


def retinanet_head_towers(head):
    """(classification tower, regression tower) of a frozen torchvision ``RetinaNetHead``, or None."""
    ch, rh = head.classification_head, head.regression_head
    out = []
    for sub, pred in ((ch, ch.cls_logits), (rh, rh.bbox_reg)):
        convs = [_conv_of(b) for b in sub.conv] if isinstance(sub.conv, torch.nn.Sequential) else [None]
        if any(c is None for c in convs) or not isinstance(pred, torch.nn.Conv2d) or pred.bias is None or pred.kernel_size != (3, 3):
            return None
        if not _frozen(pred.weight, pred.bias, *[p for c in convs for p in (c.weight, c.bias)]):
            return None
        key = _tower_key(convs, (pred.weight, pred.bias))
        cached = getattr(sub, "_hd_tower", None)
        if cached is None or cached[0] != key:
            cached = sub._hd_tower = (key, _Tower(convs, pred.weight, pred.bias, pred.weight.device))
        out.append(cached[1])
    return tuple(out)


def retinanet_head_forward(head, features, features_bf16):
    """``RetinaNetHead.forward`` (TV retinanet.py:54-57, 162-181, 254-273): {"cls_logits": [B, sum HWA, K], "bbox_regression":
    [B, sum HWA, 4]}.  The channels-last predictor output IS the (H, W, A, K) order torchvision permutes to."""
    towers = retinanet_head_towers(head)
    k = head.classification_head.num_classes
    outs = {}
    for name, tower, last in (("cls_logits", towers[0], k), ("bbox_regression", towers[1], 4)):
        preds = _tower_apply(tower, list(features_bf16), *features)
        per_level = [p[..., :tower.pred_c].reshape(p.shape[0], -1, last) for p in preds]
        outs[name] = torch.cat(per_level, dim=1)
    return outs
