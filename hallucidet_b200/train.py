"""The HalluciDet train step on the B200 modules (host-side mirror of ``EncoderDecoderLit``).

Follows the reference's train_hallucidet.py: ``forward_step`` (:161-240: U-Net on the 3x-repeated IR image,
pixel regulariser, frozen-detector loss weighted 0.1 x4, total), ``training_step`` + Lightning's
``loss.backward()`` (:243-283), ``clip_grad_value_(0.5)`` (:498-499), Adam lr 1e-4 (:56, src/config/config.py:215-219).
The two extra no-grad detector passes (:183,:186) and the plotting normalisation (:218) are optional
(``reference_extra_passes``): they do not influence the loss or the gradients.

Data parallelism (new in this build): one process per GPU, identical replicas, per-replica
BatchNorm statistics, one NCCL all-reduce(sum)/world of the hallucinator's flat fp32 gradient block per step.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .detection import Detector, install_b200_backbone
from .unet import Unet

import os as _os

FUSED_LOSS_WEIGHTS = _os.environ.get("HD_FUSED_LOSS_WEIGHTS", "1") == "1"      # weighted loss terms as one stack / multiply / sum
OVERLAP_ALLREDUCE = _os.environ.get("HD_OVERLAP_ALLREDUCE", "1") != "0"   # bucketed all-reduce underneath the backward pass
FUSED_OPTIMIZER = _os.environ.get("HD_FUSED_ADAM", "1") != "0"     # clip + Adam + bf16 re-pack as one pass (optim.FusedAdam)

LOSS_WEIGHTS = {                       # src/config/config.py:58-69
    "pixel_rgb": 0.0, "pixel_ir": 0.0, "perceptual_rgb": 0.0, "perceptual_ir": 0.0,
    "det_regression": 0.1, "det_classification": 0.1, "det_objectness": 0.1, "det_rpn_box_reg": 0.1,
    "det_bbox_ctrness": 0.1,
}


class _PixelRegulariser(torch.autograd.Function):
    """w_rgb * pixel(rgb, hal) + w_ir * pixel(ir3, hal) with the gradient produced in the same pass."""

    @staticmethod
    def forward(ctx, hal, rgb, ir, kind, w_rgb, w_ir):
        loss = torch.zeros(2, device=hal.device)
        dhal = torch.empty_like(hal)
        ops.regulariser(kind, hal.contiguous(), rgb.contiguous(), ir.contiguous(), float(w_rgb), float(w_ir), loss, dhal)
        ctx.save_for_backward(dhal)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dhal,) = ctx.saved_tensors
        # both terms normally receive the same upstream gradient (they are summed into the total loss)
        return dhal * g[0], None, None, None, None, None


def allreduce_mean_(tensors, world=None, group=None):
    """In-place mean over ranks of a list of gradient tensors (one all-reduce per tensor; pass the flat block as a
    single tensor).  Backend-agnostic: NCCL over NVLink on the GPUs, gloo in the CPU tests."""
    if not (dist.is_available() and dist.is_initialized()):
        return tensors
    world = world or dist.get_world_size(group)
    if world == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.mul_(1.0 / world)
    return tensors


def shard_batch(global_batch, rank, world):
    """Batch sharding of SURVEY.md 8(e): contiguous, equal slices of the global batch; identical replicas."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by the world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


class DevicePrefetcher:
    """Host (pinned) -> device copies of the NEXT batch on a copy stream, overlapped with the current step (what a
    DataLoader with pinned memory + non_blocking copies gives the reference's Lightning loop).  ``put`` enqueues the
    copies into one of two persistent device slots, ``get`` makes the compute stream wait for them and hands the
    device tensors over (valid until the second ``put`` after it)."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.turn = 0
        self.pending = None

    def put(self, *host_tensors):
        slot = self.slots[self.turn]
        if slot is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slot, host_tensors)):
            slot = self.slots[self.turn] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors]
        # the slot was last read two steps ago; everything enqueued so far on the compute stream is older than that read
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            for d, h in zip(slot, host_tensors):
                d.copy_(h, non_blocking=True)
        event = torch.cuda.Event()
        event.record(self.stream)
        self.pending = (slot, event)
        self.turn ^= 1

    def get(self):
        dev, event = self.pending
        self.pending = None
        torch.cuda.current_stream(self.device).wait_event(event)
        return dev


class HostScalar:
    """A device scalar on its way to the host: copied to pinned memory (non-blocking) at the point of the step where it
    exists; ``float()`` waits for THAT copy only, not for whatever the stream was given afterwards.  The train step's loss
    is final before the backward pass starts, so logging it every step does not have to wait for backward + optimizer."""

    def __init__(self, value, slot):
        slot.copy_(value.detach().reshape(()), non_blocking=True)
        self._slot = slot
        self._event = torch.cuda.Event()
        self._event.record()

    def __float__(self):
        self._event.synchronize()
        return float(self._slot)

    item = __float__


def expand_one_channel_to_output_channels(imgs, output_channels=3):
    """src/utils/utils.py:52-53 (``imgs.repeat(1, C, 1, 1)``) as a zero-copy broadcast view: same values for every reader; the
    B200 U-Net recognises the broadcast (channel stride 0) and runs its stem on the single plane."""
    if imgs.shape[1] == 1:
        return imgs.expand(-1, output_channels, -1, -1)
    return imgs.repeat(1, output_channels, 1, 1)


class HalluciDetTrainer(nn.Module):
    def __init__(self, detector_name="fasterrcnn", size=640, pixel=None, weights=None, lr=1e-4, clip_value=0.5,
                 seed=123, device="cuda", reference_extra_passes=False, use_cuda_graph=False, detector_state=None):
        super().__init__()
        torch.manual_seed(seed)
        self.detector_name = detector_name
        self.size = size
        self.pixel = pixel
        self.weights = dict(LOSS_WEIGHTS, **(weights or {}))
        self.clip_value = clip_value
        self.reference_extra_passes = reference_extra_passes
        # src/models/encoder_decoder.py:22-30
        self.encoder_decoder = Unet("resnet34", encoder_depth=5, encoder_weights=None, decoder_attention_type=None,
                                    in_channels=3, classes=3)
        self.encoder_decoder.segmentation_head[-1] = nn.Sigmoid()
        self.detector = Detector(name=detector_name, pretrained=False, n_classes=2, size=size).detector
        if detector_state is not None:
            self.detector.load_state_dict(detector_state)
        self.to(device)
        install_b200_backbone(self.detector)
        self.encoder_decoder.use_cuda_graph = use_cuda_graph
        self.detector.backbone.use_cuda_graph = use_cuda_graph
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._bucket_works = []
        if self.world > 1 and OVERLAP_ALLREDUCE:
            # gradient buckets in reverse-forward order: [head, decoder, layer4] is all-reduced while the
            # backward pass of layer3 .. stem is still running, [stem .. layer3] right after it
            self.encoder_decoder.grad_bucket_hook = self._allreduce_bucket
        # train_hallucidet.py:429-435 (Adam, lr 1e-4, default betas/eps) + clip_grad_value_(0.5) (:498-499) + the 1/world of the
        # gradient mean: one pass over the parameters that also re-packs the bf16 operands (hallucidet_b200/optim.py)
        if FUSED_OPTIMIZER:
            from .optim import FusedAdam
            self.optimizer = FusedAdam(self.encoder_decoder, lr=lr, clip_value=clip_value, grad_scale=1.0 / self.world)
        else:
            self.optimizer = torch.optim.Adam(self.encoder_decoder.parameters(), lr=lr, fused=True)

    def forward_step(self, imgs_rgb, targets_rgb, imgs_ir, targets_ir, det_seed=None):
        """train_hallucidet.py:161-240 -> dict(total, parts, hal)."""
        w = self.weights
        if imgs_ir.dtype == torch.uint8:
            # the camera plane as the dataloader reads it (src/dataloader/dataloader.py:13-73): the /255 of ToTensor and the
            # 1 -> 3 channel replication happen inside the U-Net's stem kernel; a float copy is made only if something else
            # (regulariser, the reference's extra detector passes) needs it
            hal = self.encoder_decoder.forward_ir(imgs_ir)
            need_float = self.pixel is not None or self.reference_extra_passes
            imgs_ir = imgs_ir.float().mul_(1.0 / 255.0) if need_float else None
            ir3 = expand_one_channel_to_output_channels(imgs_ir, 3) if need_float else None
        else:
            ir3 = expand_one_channel_to_output_channels(imgs_ir, 3)
            hal = self.encoder_decoder(ir3)
        if self.pixel is not None:
            reg = _PixelRegulariser.apply(hal, imgs_rgb, imgs_ir, self.pixel, w["pixel_rgb"], w["pixel_ir"])
            loss_pixel_rgb, loss_pixel_ir = reg[0], reg[1]
        else:
            loss_pixel_rgb, loss_pixel_ir = 0.0, 0.0
        if det_seed is not None:
            torch.manual_seed(det_seed)
        losses_det, detections = Detector.calculate_loss(self.detector, hal, targets_ir, train_det=False, model_name=self.detector_name)
        if self.reference_extra_passes:
            with torch.no_grad():
                for imgs, tg in ((imgs_rgb, targets_rgb), (ir3, targets_ir)):
                    _, extra = Detector.calculate_loss(self.detector, imgs, tg, train_det=False, model_name=self.detector_name)
                    if hasattr(extra, "resolve"):
                        extra.resolve()
        frcnn = "fasterrcnn" in self.detector_name
        if frcnn:
            losses_det["classification"] = losses_det["loss_classifier"]
            losses_det["bbox_regression"] = losses_det["loss_box_reg"]
        names = ("bbox_regression", "classification", "loss_objectness", "loss_rpn_box_reg")
        if frcnn and FUSED_LOSS_WEIGHTS and all(torch.is_tensor(losses_det[n]) and losses_det[n].is_cuda and losses_det[n].dim() == 0
                                                 for n in names):
            # the four weighted terms and their sum as stack -> multiply -> sum (3 launches forward, 1 backward) instead of four
            # scalar multiplies and three adds each way (train_hallucidet.py:198-207: same terms, same weights)
            wt = getattr(self, "_loss_weight_tensor", None)
            key = (w["det_regression"], w["det_classification"], w["det_objectness"], w["det_rpn_box_reg"], str(hal.device))
            if wt is None or wt[0] != key:
                wt = self._loss_weight_tensor = (key, torch.tensor(key[:4], dtype=torch.float32, device=hal.device))
            weighted = torch.stack([losses_det[n] for n in names]) * wt[1]
            loss_det_total = weighted.sum()
            for i, n in enumerate(names):
                losses_det[n] = weighted[i]
        else:
            losses_det["bbox_regression"] = losses_det["bbox_regression"] * w["det_regression"]
            losses_det["classification"] = losses_det["classification"] * w["det_classification"]
            losses_det["loss_objectness"] = losses_det["loss_objectness"] * w["det_objectness"] if frcnn else 0.0
            losses_det["loss_rpn_box_reg"] = losses_det["loss_rpn_box_reg"] * w["det_rpn_box_reg"] if frcnn else 0.0
            loss_det_total = losses_det["bbox_regression"] + losses_det["classification"] + losses_det["loss_objectness"] + \
                losses_det["loss_rpn_box_reg"]
        total = loss_det_total if self.pixel is None else loss_det_total + loss_pixel_rgb + loss_pixel_ir
        return {"total": total, "det_total": loss_det_total, "pixel_rgb": loss_pixel_rgb, "pixel_ir": loss_pixel_ir,
                "losses_det": losses_det, "hal": hal, "detections": detections}

    @torch.no_grad()
    def test_step(self, imgs_rgb, targets_rgb, imgs_ir, targets_ir, det_seed=None):
        """eval_hallucidet.py:135-161 -- eval-mode hallucination (folded BN) + detector losses / detections on it
        (the reference also evaluates the RGB and IR images; enable with ``reference_extra_passes``)."""
        self.encoder_decoder.eval()
        if imgs_ir.dtype == torch.uint8:
            hal = self.encoder_decoder.forward_ir(imgs_ir)
            ir3 = expand_one_channel_to_output_channels(imgs_ir.float().mul_(1.0 / 255.0), 3) if self.reference_extra_passes else None
        else:
            ir3 = expand_one_channel_to_output_channels(imgs_ir, 3)
            hal = self.encoder_decoder(ir3)
        if det_seed is not None:
            torch.manual_seed(det_seed)
        losses_hal, det_hal = Detector.calculate_loss(self.detector, hal, targets_ir, train_det=False, model_name=self.detector_name)
        out = {"hal": hal, "losses_hal": losses_hal, "detections_hal": det_hal}
        if self.reference_extra_passes:
            _, out["detections_rgb"] = Detector.calculate_loss(self.detector, imgs_rgb, targets_rgb, train_det=False, model_name=self.detector_name)
            _, out["detections_ir"] = Detector.calculate_loss(self.detector, ir3, targets_ir, train_det=False, model_name=self.detector_name)
        return out

    def _allreduce_bucket(self, flat_slice):
        """grad_bucket_hook of the U-Net engine: asynchronous all-reduce(sum) of one finished bucket.  NCCL orders it after the
        work already enqueued on the current stream (the backward segment that produced the bucket) and runs it on its own
        stream, next to the following segment."""
        self._bucket_works.append((dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, async_op=True), flat_slice))

    def allreduce_gradients(self, scale=True):
        """Mean (``scale``) or sum of the per-replica gradients: one all-reduce over the U-Net's flat gradient block -- or, if the
        buckets were already reduced during the backward pass (grad_bucket_hook), just the wait for them."""
        if self.world == 1:
            return
        if self._bucket_works:
            works, self._bucket_works = self._bucket_works, []
            for w, _ in works:
                w.wait()                                   # the compute stream waits for the communication stream (no host block)
            if self._flat_grad() is not None:              # p.grad are views of the reduced block
                if scale:
                    for _, t in works:
                        t.mul_(1.0 / self.world)
                return
            # (p.grad does not alias the flat block -- gradient accumulation: reduce the real gradients below)
        flat = self._flat_grad()
        tensors = [flat] if flat is not None else [p.grad for p in self.encoder_decoder.parameters() if p.grad is not None]
        if scale:
            allreduce_mean_(tensors, self.world)
        else:
            for t in tensors:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def _flat_grad(self):
        """The U-Net engine's flat fp32 gradient block, if the parameters' .grad tensors are views into it."""
        first = getattr(self, "_first_param", None)
        if first is None:
            first = self._first_param = next(iter(self.encoder_decoder.parameters()))
        eng = next((e for e in self.encoder_decoder._engines.values() if e.training), None)
        if eng is not None and first.grad is not None and first.grad.data_ptr() == eng.flat_grad.data_ptr():
            return eng.flat_grad
        return None

    def clip_gradients(self):
        """clip_grad_value_(0.5) (train_hallucidet.py:498-499): one clamp over the flat block when possible."""
        flat = self._flat_grad()
        if flat is not None:
            flat.clamp_(-self.clip_value, self.clip_value)
        else:
            torch.nn.utils.clip_grad_value_(self.encoder_decoder.parameters(), self.clip_value)

    def training_step(self, imgs_rgb, targets_rgb, imgs_ir, targets_ir, det_seed=None):
        from . import detection
        if not self.encoder_decoder.training:
            self.encoder_decoder.train()
        self.optimizer.zero_grad(set_to_none=True)
        detection.DEFER_DETECTIONS = True           # the detections are a by-product here: assemble them after the backward
        try:
            out = self.forward_step(imgs_rgb, targets_rgb, imgs_ir, targets_ir, det_seed=det_seed)
        finally:
            detection.DEFER_DETECTIONS = False
        if out["total"].is_cuda:
            # the loss as a host number (for logging): read back from here, ahead of the backward pass
            slots = getattr(self, "_loss_slots", None)
            if slots is None:
                slots = self._loss_slots = [torch.empty((), dtype=torch.float32, pin_memory=True) for _ in range(4)]
                self._loss_turn = 0
            self._loss_turn = (self._loss_turn + 1) % len(slots)
            out["total_host"] = HostScalar(out["total"].float(), slots[self._loss_turn])
        out["total"].backward()
        if FUSED_OPTIMIZER:
            self.allreduce_gradients(scale=False)        # sum over ranks; the 1/world, the clip and Adam run in the fused pass
        else:
            self.allreduce_gradients()
            self.clip_gradients()
        self.optimizer.step()
        if isinstance(out["detections"], (detection.DeferredDetections, detection.DeferredCall)):
            out["detections"] = out["detections"].resolve()
        return out
