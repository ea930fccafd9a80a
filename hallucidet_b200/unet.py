"""B200-native drop-in for the reference's hallucination U-Net.

Mirrors ``smp.Unet('resnet34', ...)`` as used at src/models/encoder_decoder.py:22-30 (reference repo):
same constructor arguments, same sub-module / ``state_dict`` names and shapes (278 entries), same
``forward(Tensor[B,3,H,W] fp32) -> Tensor[B,3,H,W] fp32`` with autograd, ``.train()/.eval()`` switching the
BatchNorm mode, ``segmentation_head[-1]`` assignable (the reference replaces it with ``nn.Sigmoid``).

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` children only HOLD parameters (fp32 masters for Adam / checkpoints);
they are never called.  All arithmetic runs in the hand-written sm_100a kernels of libhallucidet_b200.so
through ``hallucidet_b200.ops`` (NHWC bf16 activations, fp32 accumulation, fp32 BN statistics).  There is no
PyTorch/CPU fallback: a CPU tensor or a missing library raises.

Reference semantics followed: src/segmentation_models/base/model.py:24-38, encoders/resnet.py:47-65,
decoders/unet/decoder.py:7-8,38-46,111-124, base/modules.py:10-47, base/heads.py:21-27,
base/initialization.py:4-27; TV models/resnet.py:59-105 (BasicBlock).
"""
import os

import torch
import torch.nn as nn

from . import ops

BN_EPS = 1e-5


# ------------------------------------------------------------------------------------------------------
# parameter-holding module tree (names == reference state_dict)
# ------------------------------------------------------------------------------------------------------
class _Conv2dReLU(nn.Sequential):
    """Holder mirroring base/modules.py:10-47: [conv(bias=False), bn, relu]."""

    def __init__(self, cin, cout):
        super().__init__(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _Conv2dReLU(cin + cskip, cout)
        self.attention1 = nn.Identity()
        self.conv2 = _Conv2dReLU(cout, cout)
        self.attention2 = nn.Identity()


class _UnetDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        ins = [enc[0]] + list(decoder_channels[:-1])
        skips = enc[1:] + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList(_DecoderBlock(i, s, o) for i, s, o in zip(ins, skips, decoder_channels))
        self.channels = list(zip(ins, skips, decoder_channels))


def _make_encoder(name):
    from torchvision.models.resnet import ResNet, BasicBlock
    layers = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3]}
    if name not in layers:
        raise KeyError(f"Wrong encoder name `{name}`, supported encoders: {list(layers)} (B200 hot path: BasicBlock ResNets)")
    enc = ResNet(block=BasicBlock, layers=layers[name])
    del enc.fc
    del enc.avgpool
    enc.out_channels = (3, 64, 64, 128, 256, 512)
    enc.output_stride = 32
    enc.layer_spec = layers[name]
    return enc


class Unet(nn.Module):
    def __init__(self, encoder_name="resnet34", encoder_depth=5, encoder_weights=None, decoder_use_batchnorm=True,
                 decoder_channels=(256, 128, 64, 32, 16), decoder_attention_type=None, in_channels=3, classes=1,
                 activation=None, aux_params=None):
        super().__init__()
        if encoder_depth != 5 or decoder_use_batchnorm is not True or decoder_attention_type is not None \
                or in_channels != 3 or aux_params is not None or tuple(decoder_channels) != (256, 128, 64, 32, 16):
            raise NotImplementedError("hallucidet_b200.Unet implements the configuration HalluciDet uses "
                                      "(depth 5, BN decoder (256,128,64,32,16), no attention, 3 input channels)")
        if encoder_weights is not None:
            raise NotImplementedError("pretrained encoder download is not available; load a state_dict instead")
        if classes > 16:
            raise NotImplementedError("segmentation head supports up to 16 classes")
        self.encoder = _make_encoder(encoder_name)
        self.decoder = _UnetDecoder(self.encoder.out_channels, decoder_channels)
        act = {None: nn.Identity(), "identity": nn.Identity(), "sigmoid": nn.Sigmoid()}
        if activation not in act:
            raise NotImplementedError(f"activation {activation!r}")
        self.segmentation_head = nn.Sequential(nn.Conv2d(decoder_channels[-1], classes, 3, padding=1), nn.Identity(), act[activation])
        self.classification_head = None
        self.name = "u-{}".format(encoder_name)
        self.classes = classes
        self.initialize()
        self._engines = {}
        self.use_cuda_graph = False

    def initialize(self):
        """base/initialization.py:4-27 (decoder kaiming_uniform fan_in/relu, BN 1/0; head xavier_uniform, bias 0)."""
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        head = self.segmentation_head[0]
        nn.init.xavier_uniform_(head.weight)
        nn.init.constant_(head.bias, 0)

    def check_input_shape(self, x):
        """base/model.py:12-22."""
        h, w = x.shape[-2:]
        if h % 32 != 0 or w % 32 != 0:
            new_h = (h // 32 + 1) * 32 if h % 32 != 0 else h
            new_w = (w // 32 + 1) * 32 if w % 32 != 0 else w
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and width "
                               f"divisible by 32. Consider pad your images to shape ({new_h}, {new_w}).")

    def _engine(self, x):
        key = (tuple(x.shape), x.dtype, x.device.index, self.training, self.segmentation_head[0].weight.data_ptr())
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= 4:
                self._engines.clear()
            eng = _UnetEngine(self, x.shape[0], x.shape[2], x.shape[3], x.device, self.training, in_channels=x.shape[1], in_dtype=x.dtype)
            self._engines[key] = eng
        return eng

    def forward(self, x, in_scale=None):
        self.check_input_shape(x)
        if not x.is_cuda:
            raise RuntimeError("hallucidet_b200.Unet runs only on a CUDA (B200) device; there is no CPU path")
        head_act = self.segmentation_head[-1]
        if isinstance(head_act, nn.Sigmoid):
            sigmoid = True
        elif isinstance(head_act, nn.Identity):
            sigmoid = False
        else:
            raise NotImplementedError(f"segmentation head activation {type(head_act).__name__} is not implemented in the B200 path")
        if x.requires_grad and torch.is_grad_enabled():
            # HalluciDet feeds the IR image (a leaf without gradient, train_hallucidet.py:170-171); the stem computes no
            # input gradient, so asking for one must not silently return zeros
            raise NotImplementedError("hallucidet_b200.Unet does not compute the gradient w.r.t. its input image; "
                                      "detach() the input (the reference never differentiates through the IR frame)")
        if not self.training and not torch.is_grad_enabled() and x.shape[0] > EVAL_CHUNK:
            # inference over a large batch (BASELINE config 5: 64 x 1024x1280): eval-mode BatchNorm is a per-pixel affine
            # map, so the batch is run through ONE engine in chunks -- identical results, activation memory of one chunk
            return torch.cat([self.forward(c, in_scale=in_scale) for c in x.split(EVAL_CHUNK)], 0)
        if ONE_CHANNEL_STEM and x.shape[1] == 3 and x.stride(1) == 0:
            # the IR plane broadcast to three channels (expand_one_channel_to_output_channels as a zero-copy view): the stem
            # runs on the single plane with the channel-summed filter -- conv(W, [x, x, x]) == conv(sum_c W[:, c], x)
            x = x[:, :1]
        if x.shape[1] == 1:
            if x.dtype != torch.uint8:
                x = x.float()
            x = x.contiguous()
        else:
            if x.shape[1] != 3:
                raise RuntimeError(f"hallucidet_b200.Unet expects 3 input channels (or the single IR plane), got {x.shape[1]}")
            x = x.contiguous().float()
        eng = self._engine(x)
        eng.in_scale = float(in_scale) if in_scale is not None else (1.0 / 255.0 if x.dtype == torch.uint8 else 1.0)
        params = [p for _, p in eng.named_params]
        return _UnetFunction.apply(x, eng, sigmoid, *params)

    def forward_ir(self, ir, in_scale=None):
        """The hallucination of a SINGLE-channel IR batch ``[B, 1, H, W]`` -- float in [0, 1], or the uint8 camera plane
        (then ``in_scale`` defaults to 1/255, the dataloader's division: src/dataloader/dataloader.py:13-73).  Equals
        ``forward(ir.repeat(1, 3, 1, 1))`` (src/utils/utils.py:52-53 + train_hallucidet.py:170-171) without materialising
        the copies: the stem convolves the plane with the channel-summed filter."""
        if ir.dim() != 4 or ir.shape[1] != 1:
            raise RuntimeError(f"forward_ir expects [B, 1, H, W], got {tuple(ir.shape)}")
        return self.forward(ir, in_scale=in_scale)

    @torch.no_grad()
    def predict(self, x):
        if self.training:
            self.eval()
        return self.forward(x)


class _UnetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng, sigmoid, *params):
        ctx.eng = eng
        ctx.sigmoid = sigmoid
        out = eng.forward(x, sigmoid)
        ctx.generation = eng.generation
        return out

    @staticmethod
    def backward(ctx, dhal):
        eng = ctx.eng
        if not ctx.sigmoid:
            raise NotImplementedError("backward needs the sigmoid head (the HalluciDet configuration)")
        if not eng.training:
            raise NotImplementedError("backward through the eval-mode (folded BatchNorm) U-Net is not implemented")
        if ctx.generation != eng.generation:
            raise RuntimeError("hallucidet_b200.Unet: the activations saved for this backward were overwritten by a later "
                               "forward of the same module (one forward per backward is supported)")
        grads = eng.backward(dhal.contiguous().float())
        return (None, None, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------------
# execution engine: static buffers + kernel program for one (B, H, W, mode)
# ------------------------------------------------------------------------------------------------------
FUSED_STEM = os.environ.get("HD_FUSED_STEM", "1") != "0"       # halo-patch stem kernels (no patch matrix on the forward path)
ONE_CHANNEL_STEM = os.environ.get("HD_STEM_1CH", "1") != "0"   # replicated IR plane -> single-channel stem (K = 49 instead of 147)
EVAL_CHUNK = int(os.environ.get("HD_EVAL_CHUNK", "8"))       # images per engine pass in eval / no-grad mode
SIDE_STREAM_WGRAD = os.environ.get("HD_SIDE_WGRAD", "1") != "0"
# one persistent launch for the BatchNorm backward (hd_bn_bwd_fused) is opt-in: 0.38 ms less GPU time per step, but its
# full-register-file CTAs keep the side-stream work (weight gradients, detections) off the SMs: +1.4 ms per step measured
FUSED_BN_BWD = os.environ.get("HD_BN_FUSED", "0") == "1"


class _Layer:
    """One convolution (+ optional BatchNorm) of the U-Net with its packed operands and saved tensors."""

    def __init__(self, eng, name, conv, bn, cin, cout, k, stride, in_hw, cout_pad=None, packed=None):
        self.name, self.conv, self.bn = name, conv, bn
        self.cin, self.cout, self.k, self.stride = cin, cout, k, stride
        self.h_in, self.w_in = in_hw
        self.h, self.w = in_hw[0] // stride, in_hw[1] // stride
        self.cp = cout_pad or cout                       # channel count of the activation tensors
        dev = eng.device
        self.packed = packed if packed is not None else ops.PackedConv(self.cp, cin, k, dev, need_dgrad=True)
        self.z = eng.new_act(self.h, self.w, self.cp)
        if bn is not None:
            self.stats = None                            # [tiles][2][cout] per-tile partials, sized at first use
            self.mean, self.invstd, self.scale, self.shift = (torch.empty(cout, device=dev) for _ in range(4))
            self.sums = torch.zeros(2, cout, device=dev)
            self.bias = torch.empty(cout, device=dev)    # eval-mode folded shift
        self.dz = None


class _UnetEngine:
    def __init__(self, module, B, H, W, device, training, in_channels=3, in_dtype=torch.float32):
        self.m, self.B, self.H, self.W, self.device, self.training = module, B, H, W, device, training
        self.one_ch = in_channels == 1
        self.in_scale = 1.0
        self.acts = []
        enc, dec = module.encoder, module.decoder
        self.named_params = [(n, p) for n, p in module.named_parameters()]
        # flat fp32 gradient buffer (views returned to autograd, one allreduce-able block)
        sizes = [p.numel() for _, p in self.named_params]
        self.flat_grad = torch.zeros(sum(sizes), device=device)
        self.grad_views, off = {}, 0
        self.bucket_split = None                            # first element of the "late" bucket (encoder.layer4 onwards)
        for (n, p), s in zip(self.named_params, sizes):
            if self.bucket_split is None and n.startswith("encoder.layer4."):
                self.bucket_split = off
            self.grad_views[n] = self.flat_grad[off:off + s].view_as(p)
            off += s

        h2, w2 = H // 2, W // 2
        # ---- stem (7x7/2 via patches GEMM)
        self.stem_kpad = ops.STEM1_KPAD if self.one_ch else ops.STEM_KPAD
        self.stem = _Layer(self, "encoder.conv1", enc.conv1, enc.bn1, 1 if self.one_ch else 3, 64, 7, 2, (H, W),
                           packed=ops.PackedConv(64, 1 if self.one_ch else 3, 7, device, need_dgrad=False, k_pad=self.stem_kpad))
        self.patches = torch.empty(1, 1, B * h2 * w2, self.stem_kpad, dtype=torch.bfloat16, device=device)
        if self.one_ch:
            self.stem_w1 = torch.empty(64, 1, 7, 7, device=device)      # channel-summed stem filter (re-summed at every pack)
            self.stem_g1 = torch.empty(64, 1, 7, 7, device=device)      # its gradient (== the gradient of every input channel's filter)
        self.a_stem = self.new_act(h2, w2, 64)
        self.p0 = self.new_act(h2 // 2, w2 // 2, 64)
        self.p0_idx = torch.empty(B, h2 // 2, w2 // 2, 64, dtype=torch.uint8, device=device)   # max-pool arg-max positions
        # ---- encoder layers
        self.blocks = []
        hw, cin = (h2 // 2, w2 // 2), 64
        for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), enc.layer_spec), start=1):
            layer = getattr(enc, f"layer{li}")
            for b in range(nblk):
                blk = layer[b]
                stride = 2 if (li > 1 and b == 0) else 1
                pfx = f"encoder.layer{li}.{b}"
                c1 = _Layer(self, pfx + ".conv1", blk.conv1, blk.bn1, cin, planes, 3, stride, hw)
                ohw = (hw[0] // stride, hw[1] // stride)
                c2 = _Layer(self, pfx + ".conv2", blk.conv2, blk.bn2, planes, planes, 3, 1, ohw)
                cd = None
                if blk.downsample is not None:
                    cd = _Layer(self, pfx + ".downsample.0", blk.downsample[0], blk.downsample[1], cin, planes, 1, stride, hw)
                    cd.bn_name = pfx + ".downsample.1"
                c1.bn_name, c2.bn_name = pfx + ".bn1", pfx + ".bn2"
                a1 = self.new_act(ohw[0], ohw[1], planes)
                out = self.new_act(ohw[0], ohw[1], planes)
                self.blocks.append(dict(c1=c1, c2=c2, cd=cd, a1=a1, out=out, end_of_layer=(b == nblk - 1), li=li))
                hw, cin = ohw, planes
        # ---- decoder
        self.dblocks = []
        x_hw, x_c = hw, cin
        skips_c = [256, 128, 64, 64, 0]
        for i, (ci, cs, co) in enumerate(dec.channels):
            blk = dec.blocks[i]
            ohw = (x_hw[0] * 2, x_hw[1] * 2)
            up = self.new_act(ohw[0], ohw[1], ci)
            c1 = _Layer(self, f"decoder.blocks.{i}.conv1.0", blk.conv1[0], blk.conv1[1], ci + cs, co, 3, 1, ohw)
            c2 = _Layer(self, f"decoder.blocks.{i}.conv2.0", blk.conv2[0], blk.conv2[1], co, co, 3, 1, ohw)
            c1.bn_name, c2.bn_name = f"decoder.blocks.{i}.conv1.1", f"decoder.blocks.{i}.conv2.1"
            a1 = self.new_act(ohw[0], ohw[1], co)
            a2 = self.new_act(ohw[0], ohw[1], co)
            self.dblocks.append(dict(c1=c1, c2=c2, up=up, a1=a1, a2=a2, cskip=cs, cin=ci))
            x_hw, x_c = ohw, co
        # ---- head (cout padded to 16 channels)
        self.head = _Layer(self, "segmentation_head.0", module.segmentation_head[0], None, 16, module.classes, 3, 1, (H, W), cout_pad=16)
        self.head_w_pad = torch.zeros(16, 16, 3, 3, device=device)
        self.head_bias_pad = torch.zeros(16, device=device)
        self.hal = torch.empty(B, module.classes, H, W, device=device)
        self.dlogits = self.new_act(H, W, 16)
        self.stem.bn_name = "encoder.bn1"
        self.all_layers = [self.stem] + [l for b in self.blocks for l in (b["c1"], b["c2"], b["cd"]) if l is not None] + \
                          [l for d in self.dblocks for l in (d["c1"], d["c2"])] + [self.head]
        # packed weight-gradient accumulators (one flat buffer, zeroed once per backward)
        tot = 0
        for l in self.all_layers:
            l.dw_off = tot
            l.dw_rows = l.cp
            l.dw_cols = l.k * l.k * l.cin if l is not self.stem else self.stem_kpad
            tot += l.dw_rows * l.dw_cols
        # ... followed by the BatchNorm backward sums of every layer, so that the same fill zeroes them too
        bn_layers = [l for l in self.all_layers if getattr(l, "sums", None) is not None]
        for l in bn_layers:
            l.sums_off = tot
            tot += (l.sums.numel() + 3) // 4 * 4
        self.bn_barrier = torch.zeros(2, dtype=torch.int32, device=device)     # grid barrier of the fused BatchNorm backward
        self.bn_counters = torch.zeros(len(bn_layers), dtype=torch.int32, device=device)   # "last CTA" counters of the fused finalize
        for i, l in enumerate(bn_layers):
            l.counter = self.bn_counters[i:i + 1]
        self.dw_flat = torch.zeros(tot, device=device)
        for l in self.all_layers:
            l.dw = self.dw_flat[l.dw_off:l.dw_off + l.dw_rows * l.dw_cols]
        for l in bn_layers:
            l.sums = self.dw_flat[l.sums_off:l.sums_off + l.sums.numel()].view_as(l.sums)
        self.grad_bufs = {}
        self.side_stream, self.side_used = None, False
        self.pack_tab = self.unpack_tab = None
        self._packed_key = None
        self._bn_keys = None
        self.graphs = {}
        self.sigmoid = None
        self.generation = 0
        self.x_in = torch.empty(B, in_channels, H, W, dtype=in_dtype, device=device)
        self.dhal_in = torch.empty(B, module.classes, H, W, device=device)
        self.nbt = [m.num_batches_tracked for m in module.modules() if isinstance(m, nn.BatchNorm2d)]

    # ---- helpers -------------------------------------------------------------------------------------
    def new_act(self, h, w, c):
        t = torch.empty(self.B, h, w, c, dtype=torch.bfloat16, device=self.device)
        self.acts.append(t)
        return t

    def gbuf(self, key, like):
        t = self.grad_bufs.get(key)
        if t is None:
            t = torch.empty_like(like)
            self.grad_bufs[key] = t
        return t

    def _pack_sources(self):
        def src(l):
            if l is self.head:
                return self.head_w_pad
            if l is self.stem and self.one_ch:
                return self.stem_w1
            return l.conv.weight.detach()
        return [src(l) for l in self.all_layers]

    def _check_tables(self):
        """(Re)build the device descriptor tables of the one-launch weight packing / gradient unpacking.  They hold raw
        parameter pointers, so a parameter that was re-allocated (``.to()``, ``load_state_dict(assign=True)``) invalidates
        them -- and any captured CUDA graph."""
        if not self.training:
            return
        srcs = self._pack_sources()
        bn_keys = [l.bn.weight.data_ptr() for l in self.all_layers if l.bn is not None]
        if self.pack_tab is not None and self.pack_tab.keys == [w.data_ptr() for w in srcs] and bn_keys == self._bn_keys:
            return
        self._bn_keys = bn_keys
        assert all(w.is_contiguous() for w in srcs)
        self.pack_tab = ops.pack_table([l.packed for l in self.all_layers], srcs, self.device)
        gv = self.grad_views
        def late(l):                                         # head, decoder, encoder.layer4: see _backward_late
            return not l.name.startswith("encoder.") or l.name.startswith("encoder.layer4.")
        convs = [l for l in self.all_layers if l is not self.stem]
        ent = lambda l: (l.dw, gv[l.name + ".weight"], l.cout, l.cin, l.k, l.cin, l.k * l.k * l.cin)
        early = [ent(l) for l in convs if not late(l)]
        if self.one_ch:
            early.append((self.stem.dw, self.stem_g1, 64, 1, 7, 1, self.stem_kpad))
        else:
            early.append((self.stem.dw, gv["encoder.conv1.weight"], 64, 3, 7, 3, self.stem_kpad))
        self.unpack_tab_late = ops.unpack_table([ent(l) for l in convs if late(l)], self.device)
        self.unpack_tab_early = ops.unpack_table(early, self.device)
        self.unpack_tab = self.unpack_tab_early
        self._packed_key = None
        self.graphs.clear()

    def _version_key(self):
        """Changes whenever a parameter (or, in eval mode, a BatchNorm buffer) may have changed: in-place torch updates bump
        ``_version``; the fused optimizer (raw-pointer writes) bumps ``module._param_generation``; the BatchNorm kernels
        (raw-pointer writes to the running statistics) are followed by the ``num_batches_tracked`` increment."""
        key = [getattr(self.m, "_param_generation", 0)] + [p._version for _, p in self.named_params]
        if not self.training:
            key += [b._version for b in self.m.buffers()]
        return key

    def _pack_prologue(self):
        """Small staging copies the pack launch reads: the head weight / bias padded to 16 channels, the channel-summed stem filter."""
        if self.one_ch:
            torch.sum(self.stem.conv.weight.detach(), dim=1, keepdim=True, out=self.stem_w1)
        hd = self.head
        self.head_w_pad[:hd.cout].copy_(hd.conv.weight.detach())
        self.head_bias_pad[:hd.cout].copy_(hd.conv.bias.detach())

    def mark_weights_packed(self):
        """The bf16 operands of THIS engine match the fp32 masters (called by hallucidet_b200.optim.FusedAdam, whose kernel
        re-packs while it updates): the next forward skips the pack launch.  Other engines of the module (eval mode) are told
        that the parameters moved."""
        self.m._param_generation = getattr(self.m, "_param_generation", 0) + 1
        self._packed_key = self._version_key()

    def _pack_weights(self):
        if self.training:                                 # one launch for all layers (fp32 masters -> bf16 GEMM operands)
            self._pack_prologue()
            ops.pack_conv_weights(self.pack_tab)
            return
        if self.one_ch:
            torch.sum(self.stem.conv.weight.detach(), dim=1, keepdim=True, out=self.stem_w1)
        for l in self.all_layers:
            if l is self.head:
                self.head_w_pad[:l.cout].copy_(l.conv.weight.detach())
                self.head_bias_pad[:l.cout].copy_(l.conv.bias.detach())
                l.packed.pack(self.head_w_pad)
            elif l.bn is None:
                l.packed.pack(l.conv.weight.detach().contiguous())
            else:                                         # eval: fold running statistics (TV ops/misc.py-style affine)
                bn = l.bn
                scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
                l.bias.copy_(bn.bias.detach() - bn.running_mean * scale)
                w = self.stem_w1 if (l is self.stem and self.one_ch) else l.conv.weight.detach().contiguous()
                l.packed.pack(w, scale.contiguous())
                if l is self.stem:
                    self.stem_fold_scale = scale.contiguous()

    def _bn_fin(self, l, count):
        """The layer's hd_bn_fin descriptor (raw pointers to its BatchNorm parameters / buffers; rebuilt if they moved)."""
        bn = l.bn
        key = (bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr())
        fin = getattr(l, "fin", None)
        if fin is None or fin._key != key:
            fin = l.fin = ops.bn_fin(count, bn.weight.detach(), bn.bias.detach(), bn.eps,
                                     bn.momentum if bn.momentum is not None else 0.1, bn.running_mean, bn.running_var,
                                     l.mean, l.invstd, l.scale, l.shift, l.counter)
        return fin

    def _conv_bn(self, l, x0, a_out, x1=None, relu=True, res=None, res_layer=None, apply=True):
        """train: z = conv(x) (+stats) -> finalize -> a_out = relu(bn(z) + res).   eval: fused epilogue."""
        if self.training:
            if l.stats is None:
                l.stats = torch.zeros(ops.conv_fwd_tiles(x0, l.k, l.stride, cout=None if x1 is not None else l.cout), 2, l.cout, device=self.device)
            # statistics + BatchNorm finalize (scale / shift / running statistics) in the tail of the convolution itself
            ops.conv_fwd(ops.conv_args(x0, l.z, l.packed.w_fwd, k=l.k, stride=l.stride, x1=x1, stats=l.stats,
                                       bn_fin=self._bn_fin(l, self.B * l.h * l.w)))
            if apply:
                if res_layer is not None:
                    ops.bn_apply(l.z, l.scale, l.shift, a_out, relu=relu, res=res_layer.z, res_scale=res_layer.scale, res_shift=res_layer.shift)
                else:
                    ops.bn_apply(l.z, l.scale, l.shift, a_out, relu=relu, res=res)
        else:
            add = res_layer.z if res_layer is not None else res
            ops.conv_fwd(ops.conv_args(x0, a_out if apply else l.z, l.packed.w_fwd, k=l.k, stride=l.stride, x1=x1, bias=l.bias,
                                       add=add, relu=relu and apply))

    # ---- forward ---------------------------------------------------------------------------------------
    def _run(self, kind, fn):
        """Run a kernel program eagerly, or (module.use_cuda_graph) warm up once, capture once, then replay."""
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

    def forward(self, x, sigmoid):
        if self.sigmoid is not None and self.sigmoid != sigmoid:
            self.graphs.clear()
        self.sigmoid = sigmoid
        self._check_tables()
        key = self._version_key()
        if key != self._packed_key:                          # (skipped when the fused optimizer has just re-packed, and in
            self._pack_weights()                             #  eval mode while the weights are unchanged)
            self._packed_key = key
        self.x_in.copy_(x)
        self._run("fwd", self._forward_impl)
        self.generation += 1
        return self.hal.clone()

    def _grads_alias_flat(self, params):
        """True if some parameter's ``.grad`` lives inside ``flat_grad`` (autograd kept the views returned by an earlier
        backward: ``zero_grad(set_to_none=False)``, gradient accumulation, Lightning's ``accumulate_grad_batches``)."""
        lo = self.flat_grad.data_ptr()
        hi = lo + self.flat_grad.numel() * 4
        return any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in params)

    def backward(self, dhal):
        self.dhal_in.copy_(dhal)
        params = [p for _, p in self.named_params]
        # The kernels OVERWRITE flat_grad.  If accumulated gradients live in it, they are set aside, the new gradients are
        # handed to autograd as a separate tensor, and AccumulateGrad adds them in place: p.grad (still a view of
        # flat_grad) ends up as old + new, as with any other module.
        saved = self.flat_grad.clone() if self._grads_alias_flat(params) else None
        hook = getattr(self.m, "grad_bucket_hook", None)
        if hook is not None and saved is None:
            # data parallelism: the two gradient buckets are handed to the hook (an asynchronous all-reduce) as soon as they are
            # final -- the first one while the rest of the backward pass still runs (reverse-forward bucket order)
            self._run("bwd_late", self._backward_late)
            hook(self.flat_grad[self.bucket_split:])
            self._run("bwd_early", self._backward_early)
            hook(self.flat_grad[:self.bucket_split])
        else:
            self._run("bwd", self._backward_impl)
        alias_ok = saved is None and all(p.grad is None for p in params)
        src = self.flat_grad if alias_ok else self.flat_grad.clone()
        if saved is not None:
            self.flat_grad.copy_(saved)
        out, off = [], 0
        for p in params:
            n = p.numel()
            out.append(src[off:off + n].view_as(p) if p.requires_grad else None)
            off += n
        return out

    def _stem_patches(self):
        if self.one_ch:
            ops.stem_im2col_1ch(self.x_in, self.patches, scale=self.in_scale, k_pad=self.stem_kpad)
        else:
            ops.stem_im2col(self.x_in, self.patches)

    def _forward_impl(self):
        sigmoid = self.sigmoid
        x = self.x_in
        st = self.stem
        w_stem = st.conv.weight.detach()
        zs = st.z.view(1, 1, -1, 64)
        a_stem_flat = self.a_stem.view(1, 1, -1, 64)
        if FUSED_STEM:
            # fused halo-patch stem (csrc/stem_conv.cu): no patch matrix on the forward path.  The patches are only needed by the
            # stem's weight gradient at the very end of the backward pass: they are built on the side stream.
            if self.training:
                self._on_side(self._stem_patches)
                if st.stats is None:
                    st.stats = torch.zeros(ops.stem_fwd_rows(x), 2, 64, device=self.device)
                ops.stem_fwd(x, w_stem, st.z, x_scale=self.in_scale, stats=st.stats,
                             bn_fin=self._bn_fin(st, self.B * st.h * st.w))
                ops.bn_apply(zs, st.scale, st.shift, a_stem_flat, relu=True)
            else:
                ops.stem_fwd(x, w_stem, self.a_stem, x_scale=self.in_scale, w_scale=self.stem_fold_scale, bias=st.bias, relu=True)
        else:
            self._stem_patches()
            stem_k = 49 if self.one_ch else 147
            if self.training:
                if st.stats is None:
                    st.stats = torch.zeros(ops.conv_fwd_tiles(self.patches, 1, 1), 2, 64, device=self.device)
                ops.conv_fwd(ops.conv_args(self.patches, zs, st.packed.w_fwd, k=1, stats=st.stats, algo_cin=stem_k,
                                           bn_fin=self._bn_fin(st, zs.shape[2])))
                ops.bn_apply(zs, st.scale, st.shift, a_stem_flat, relu=True)
            else:
                ops.conv_fwd(ops.conv_args(self.patches, a_stem_flat, st.packed.w_fwd, k=1, bias=st.bias, relu=True, algo_cin=stem_k))
        ops.maxpool_fwd(self.a_stem, self.p0, idx=self.p0_idx if self.training else None)
        x_in = self.p0
        feats = {1: self.a_stem}
        for blk in self.blocks:
            c1, c2, cd = blk["c1"], blk["c2"], blk["cd"]
            self._conv_bn(c1, x_in, blk["a1"])
            if cd is not None:
                self._conv_bn(cd, x_in, None, apply=False, relu=False)
                self._conv_bn(c2, blk["a1"], blk["out"], res_layer=cd)
            else:
                self._conv_bn(c2, blk["a1"], blk["out"], res=x_in)
            blk["x_in"] = x_in
            x_in = blk["out"]
            if blk["end_of_layer"]:
                feats[blk["li"] + 1] = x_in
        skips = [feats[4], feats[3], feats[2], feats[1], None]
        xd = feats[5]
        for i, d in enumerate(self.dblocks):
            ops.upsample2x_fwd(xd, d["up"])
            d["skip"] = skips[i]
            self._conv_bn(d["c1"], d["up"], d["a1"], x1=skips[i])
            self._conv_bn(d["c2"], d["a1"], d["a2"])
            xd = d["a2"]
        hd = self.head
        ops.conv_fwd(ops.conv_args(xd, hd.z, hd.packed.w_fwd, k=3, bias=self.head_bias_pad, sigmoid=sigmoid, out_f32=self.hal,
                                   out_f32_channels=hd.cout, store_bf16=False, algo_cout=hd.cout))
        self.head_in = xd
        self._join_side()
        if self.training:
            torch._foreach_add_(self.nbt, 1)

    # ---- backward --------------------------------------------------------------------------------------
    def _bn_bwd(self, l, g, y_relu, z=None, g_out=None, direct_relu=False):
        """BatchNorm backward for layer l given g = dL/d(ReLU output); the ReLU mask comes from y_relu (residual blocks:
        the block output) or, when the ReLU follows this BN directly, is recomputed from z (one tensor less to read)."""
        z = l.z if z is None else z                   # (l.sums was zeroed with dw_flat at the start of the backward)
        rs, rb = (l.scale, l.shift) if direct_relu else (None, None)
        yr = None if direct_relu else y_relu
        dz = self.gbuf(("dz", l.name), z)
        if FUSED_BN_BWD:                              # one persistent launch: reduce, grid barrier, apply
            ops.bn_bwd_fused(g, yr, z, l.mean, l.invstd, l.bn.weight.detach(), l.sums, dz, self.bn_barrier, g_out,
                             self.grad_views[l.bn_name + ".weight"], self.grad_views[l.bn_name + ".bias"],
                             relu_scale=rs, relu_shift=rb)
            return dz
        ops.bn_bwd_reduce(g, yr, z, l.mean, l.invstd, l.sums, relu_scale=rs, relu_shift=rb)
        ops.bn_bwd_apply(g, yr, z, l.mean, l.invstd, l.bn.weight.detach(), l.sums, dz, g_out,
                         self.grad_views[l.bn_name + ".weight"], self.grad_views[l.bn_name + ".bias"],
                         relu_scale=rs, relu_shift=rb)
        return dz

    def _on_side(self, fn):
        """Run ``fn`` (launches only) on the engine's side stream, ordered after everything enqueued so far on the current
        stream.  The weight gradients are leaves of the backward dependency graph (nothing reads them before the end), so
        they overlap the dgrad -> BN-backward critical path and fill the SMs the small layers leave idle.  Works eagerly
        and under CUDA-graph capture (fork here, join at the end of ``_backward_impl``)."""
        if not SIDE_STREAM_WGRAD:
            fn()
            return
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream(device=self.device)
        self.side_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side_stream):
            fn()
        self.side_used = True

    def _wgrad(self, l, x0, dz, x1=None):
        self._on_side(lambda: ops.conv_wgrad(ops.conv_args(x0, dz, k=l.k, stride=l.stride, x1=x1, dw=l.dw, algo_cout=l.cout)))

    def _join_side(self):
        if self.side_used:                                   # join: every weight-gradient accumulator issued so far is complete
            torch.cuda.current_stream().wait_stream(self.side_stream)
            self.side_used = False

    def _backward_impl(self):
        self._backward_late()
        self._backward_early()

    def _backward_late(self):
        """Head, decoder and encoder.layer4: the LAST parameters of the model = a contiguous suffix of the flat gradient
        block (24.4 M parameters: 16.3 M here).  Under data parallelism this bucket is all-reduced while ``_backward_early``
        (layer3 .. stem) still computes."""
        dhal = self.dhal_in
        self.dw_flat.zero_()
        hd = self.head
        gv = self.grad_views
        gv["segmentation_head.0.bias"].zero_()
        ops.sigmoid_bwd_pack(dhal, self.hal, self.dlogits, gv["segmentation_head.0.bias"])
        self._wgrad(hd, self.head_in, self.dlogits)
        g = self.gbuf(("g", "head_in"), self.head_in)
        ops.conv_dgrad(ops.conv_args(self.dlogits, g, hd.packed.w_dgrad, k=3, algo_cout=hd.cout))
        # ---- decoder
        skip_grads = {}
        for i in reversed(range(len(self.dblocks))):
            d = self.dblocks[i]
            c1, c2 = d["c1"], d["c2"]
            dz2 = self._bn_bwd(c2, g, d["a2"], direct_relu=True)
            self._wgrad(c2, d["a1"], dz2)
            g_a1 = self.gbuf(("g", c2.name), d["a1"])
            ops.conv_dgrad(ops.conv_args(dz2, g_a1, c2.packed.w_dgrad, k=3))
            dz1 = self._bn_bwd(c1, g_a1, d["a1"], direct_relu=True)
            self._wgrad(c1, d["up"], dz1, x1=d["skip"])
            g_up = self.gbuf(("gup", i), d["up"])
            g_skip = self.gbuf(("gskip", i), d["skip"]) if d["skip"] is not None else None
            ops.conv_dgrad(ops.conv_args(dz1, g_up, c1.packed.w_dgrad, k=3, y1=g_skip))
            skip_grads[i] = g_skip
            x_small_like = self.dblocks[i - 1]["a2"] if i > 0 else self.blocks[-1]["out"]
            g = self.gbuf(("gdown", i), x_small_like)
            ops.upsample2x_bwd(g_up, g)
        # skips: decoder block 0 <- f4 (layer3 out), 1 <- f3 (layer2 out), 2 <- f2 (layer1 out), 3 <- f1 (stem)
        skip_for_layer = {3: skip_grads[0], 2: skip_grads[1], 1: skip_grads[2]}
        # ---- encoder.layer4
        self._bw_state = (g, skip_grads, skip_for_layer)
        g = self._backward_blocks([b for b in self.blocks if b["li"] == 4], g, skip_for_layer)
        self._bw_state = (g, skip_grads, skip_for_layer)
        self._join_side()
        ops.unpack_wgrads(self.unpack_tab_late)

    def _backward_early(self):
        """encoder.layer3 .. stem (the prefix of the flat gradient block)."""
        g, skip_grads, skip_for_layer = self._bw_state
        g = self._backward_blocks([b for b in self.blocks if b["li"] < 4], g, skip_for_layer)
        self._backward_stem(g, skip_grads)

    def _backward_blocks(self, blocks, g, skip_for_layer):
        for blk in reversed(blocks):
            c1, c2, cd = blk["c1"], blk["c2"], blk["cd"]
            g_masked = self.gbuf(("gm", c2.name), blk["out"]) if cd is None else None
            dz2 = self._bn_bwd(c2, g, blk["out"], g_out=g_masked)
            dzd = self._bn_bwd(cd, g, blk["out"]) if cd is not None else None
            self._wgrad(c2, blk["a1"], dz2)
            g_a1 = self.gbuf(("g", c2.name), blk["a1"])
            ops.conv_dgrad(ops.conv_args(dz2, g_a1, c2.packed.w_dgrad, k=3))
            dz1 = self._bn_bwd(c1, g_a1, blk["a1"], direct_relu=True)
            x_in = blk["x_in"]
            self._wgrad(c1, x_in, dz1)
            g_x = self.gbuf(("gx", c1.name), x_in)
            if cd is None:
                ops.conv_dgrad(ops.conv_args(dz1, g_x, c1.packed.w_dgrad, k=3, stride=c1.stride, add=g_masked))
            else:
                # input of a down-sampling block is the previous layer's output, which also feeds a decoder skip
                ops.conv_dgrad(ops.conv_args(dz1, g_x, c1.packed.w_dgrad, k=3, stride=c1.stride, add=skip_for_layer[blk["li"] - 1]))
                self._wgrad(cd, x_in, dzd)
                ops.conv_dgrad(ops.conv_args(dzd, g_x, cd.packed.w_dgrad, k=1, stride=cd.stride, add=g_x))
            g = g_x
        return g

    def _backward_stem(self, g, skip_grads):
        # ---- stem: max-pool backward (+ decoder skip of f1), BN backward, weight gradient through the patch GEMM
        st = self.stem
        g_stem = self.gbuf(("g", "stem"), self.a_stem)
        ops.maxpool_bwd(self.a_stem, self.p0, g, g_stem, add=skip_grads[3], idx=self.p0_idx)
        zs = st.z.view(1, 1, -1, 64)
        dzs = self._bn_bwd(st, g_stem.view(1, 1, -1, 64), self.a_stem.view(1, 1, -1, 64), z=zs, direct_relu=True)
        self._on_side(lambda: ops.conv_wgrad(ops.conv_args(self.patches, dzs, k=1, dw=st.dw, algo_cin=49 if self.one_ch else 147)))
        self._join_side()
        ops.unpack_wgrads(self.unpack_tab_early)             # packed fp32 accumulators -> the flat OIHW gradient block
        if self.one_ch:                                      # x_c identical for all c  =>  dL/dW[:, c] identical: broadcast
            self.grad_views["encoder.conv1.weight"].copy_(self.stem_g1.expand(-1, 3, -1, -1))
