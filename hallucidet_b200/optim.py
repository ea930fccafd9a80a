"""The optimizer tail of the HalluciDet train step as ONE pass over the hallucinator's parameters.

Reference: ``torch.optim.Adam(lr=1e-4)`` (train_hallucidet.py:429-435, src/config/config.py:215-219) after
``clip_grad_value_(0.5)`` (train_hallucidet.py:498-499); under data parallelism the gradients are first averaged.  With the
stock pieces that is: scale (1/world) -> clamp -> multi-tensor Adam (reads p, g, m, v; writes p, m, v) -> and, for this
build, the fp32 -> bf16 re-pack of every conv weight at the next forward: four passes over 24.4 M parameters.

``FusedAdam.step()`` issues two launches instead:
  * ``hd_adam_multi``            element-wise clip + Adam for the small tensors (BatchNorm weights / biases, head, stem), then
  * ``hd_adam_pack_conv_weights`` clip + Adam + bf16 re-pack of every conv layer: the master weight, its gradient and both
    moments are read once, the new master weight and moments are written once, and both bf16 GEMM operand layouts are
    produced from the new value in the same kernel (the next forward does not re-pack).

Drop-in for ``torch.optim.Adam`` over ``Unet.parameters()``: same constructor arguments (weight_decay / amsgrad / maximize
must be off, as in the reference), ``param_groups`` (an LR scheduler such as the reference's ReduceLROnPlateau works),
``state_dict()`` with ``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter.  There is no PyTorch fallback: the parameters
must belong to a ``hallucidet_b200.unet.Unet`` on a CUDA device.
"""
import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, unet, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, maximize=False,
                 clip_value=None, grad_scale=1.0):
        if weight_decay != 0 or amsgrad or maximize:
            raise NotImplementedError("FusedAdam implements the reference's Adam: weight_decay=0, amsgrad=False, maximize=False")
        from .unet import Unet
        if not isinstance(unet, Unet):
            raise TypeError("FusedAdam(unet, ...): the first argument is the hallucidet_b200.unet.Unet whose parameters it updates")
        self.unet = unet
        self.clip_value = clip_value
        self.grad_scale = float(grad_scale)
        params = [p for p in unet.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self._steps = 0
        self._tables = None

    # ---- state in torch.optim.Adam's layout (flat moment buffers, per-parameter views) -----------------------------------
    def _init_state(self):
        params = self.param_groups[0]["params"]
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self._flat_m = torch.zeros(n, device=dev)
        self._flat_v = torch.zeros(n, device=dev)
        self._step_t = torch.zeros((), dtype=torch.float32)
        off = 0
        for p in params:
            st = self.state[p]
            st["step"] = self._step_t
            st["exp_avg"] = self._flat_m[off:off + p.numel()].view_as(p)
            st["exp_avg_sq"] = self._flat_v[off:off + p.numel()].view_as(p)
            off += p.numel()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        params = self.param_groups[0]["params"]
        if params and "exp_avg" in self.state.get(params[0], {}):
            # re-home the loaded moments in the flat buffers the kernels address
            loaded = [(self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"], self.state[p]["step"]) for p in params]
            self._steps = int(float(loaded[0][2]))
            self._init_state()
            for p, (m, v, _) in zip(params, loaded):
                self.state[p]["exp_avg"].copy_(m)
                self.state[p]["exp_avg_sq"].copy_(v)
            self._step_t.fill_(self._steps)
        self._tables = None

    # ---- descriptor tables ------------------------------------------------------------------------------------------------
    def _engine(self):
        eng = next((e for e in self.unet._engines.values() if e.training), None)
        if eng is None:
            raise RuntimeError("FusedAdam.step(): the U-Net has not run a training-mode forward/backward yet")
        return eng

    def _build_tables(self, eng):
        params = self.param_groups[0]["params"]
        if not self.state or "exp_avg" not in self.state.get(params[0], {}):
            self._init_state()
        eng._check_tables()
        name_of = {id(p): n for n, p in eng.named_params}
        gv = eng.grad_views
        conv_of = {id(l.conv.weight): l for l in eng.all_layers}
        adam, small = [], []
        fused_ids = set()
        for l in eng.all_layers:
            w = l.conv.weight
            tiled = l is not eng.head and l is not eng.stem and ops.pack_tiled_ok(l.packed) and w.requires_grad
            if tiled:
                st = self.state[w]
                adam.append((gv[name_of[id(w)]], st["exp_avg"], st["exp_avg_sq"]))
                fused_ids.add(id(w))
            else:
                adam.append(None)
        for p in params:
            if id(p) in fused_ids:
                continue
            st = self.state[p]
            small.append((p.detach(), gv[name_of[id(p)]], st["exp_avg"], st["exp_avg_sq"]))
        dev = params[0].device
        pack_tab = ops.pack_table([l.packed for l in eng.all_layers], eng._pack_sources(), dev, adam=adam)
        small_tab = ops.adam_table(small, dev) if small else None
        self._tables = (eng, pack_tab, small_tab, [p.data_ptr() for p in params], eng.flat_grad.data_ptr())

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdam.step(closure) is not supported")
        eng = self._engine()
        params = self.param_groups[0]["params"]
        if (self._tables is None or self._tables[0] is not eng or self._tables[3] != [p.data_ptr() for p in params]
                or self._tables[4] != eng.flat_grad.data_ptr()):
            self._build_tables(eng)
        _, pack_tab, small_tab, _, _ = self._tables
        # the kernels read the gradients from the engine's flat block: that is where backward leaves them (p.grad are views of
        # it); gradients that live elsewhere (accumulated by hand, produced by another path) are gathered into it first
        lo, hi = eng.flat_grad.data_ptr(), eng.flat_grad.data_ptr() + eng.flat_grad.numel() * 4
        name_of = {id(p): n for n, p in eng.named_params}
        for p in params:
            if p.grad is None:
                eng.grad_views[name_of[id(p)]].zero_()
            elif not (lo <= p.grad.data_ptr() < hi):
                eng.grad_views[name_of[id(p)]].copy_(p.grad)
        g = self.param_groups[0]
        self._steps += 1
        args = ops.adam_args(g["lr"], g["betas"][0], g["betas"][1], g["eps"], self._steps, grad_scale=self.grad_scale,
                             clip=self.clip_value if self.clip_value else 0.0)
        if small_tab is not None:
            ops.adam_multi(small_tab, args)
        eng._pack_prologue()                      # head padding / channel-summed stem filter from the just-updated small tensors
        ops.adam_pack_conv_weights(pack_tab, args)
        eng.mark_weights_packed()
        self._step_t.fill_(self._steps)
        return None
