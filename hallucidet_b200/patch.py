"""Bind the reference checkout (heitorrapela/HalluciDet) to the B200 hot path without editing its files.

The reference resolves its models through module-level names at CALL time:

  * the hallucination U-Net through ``import src.segmentation_models as smp`` + ``smp.Unet(...)``
    (src/models/encoder_decoder.py:5,22);
  * the detector through ``Detector.__init__`` -> ``Detector.select_detector`` and
    ``Detector.change_generalized_transform`` -> ``CustomGeneralizedRCNNTransform`` (src/models/detector.py:39-48,94-101,122-141);
  * the loss forwards through ``eval_forward_fasterrcnn`` / ``eval_forward_retinanet`` imported into
    ``src.models.detector`` (src/models/detector.py:7-8,104-117).

``apply()`` rebinds those names:

  1. ``src.segmentation_models.Unet`` (and ``...decoders.unet.Unet`` / ``...decoders.unet.model.Unet``)
     -> ``hallucidet_b200.unet.Unet``.  If the reference's ``segmentation_models`` package cannot be imported (its
     ``encoders/__init__.py`` pulls timm / pretrainedmodels / efficientnet_pytorch), a stand-in package that exposes
     ``Unet`` only is registered under that name, so ``src.models.encoder_decoder`` imports cleanly.
  2. ``src.models.detector.CustomGeneralizedRCNNTransform`` -> the fused batched transform.
  3. ``Detector.__init__`` is wrapped: once the reference has built and re-headed the torchvision detector, it is frozen
     and its ``.backbone`` replaced by ``FrozenBackbone`` (dgrad only).  The B200 backbone keeps ``body`` / ``fpn`` as
     sub-modules under their torchvision names, so the checkpoint that ``load_from_checkpoint`` loads AFTERWARDS
     (train_hallucidet.py:107-115) lands in the same keys, and re-folds its BatchNorm constants after every load.
  4. optionally (``fast_tail=True``) ``eval_forward_fasterrcnn`` / ``eval_forward_retinanet`` -> the batched, sync-free
     restatements of ``hallucidet_b200.detection`` (identical losses and detections).

``python -m hallucidet_b200.run train_hallucidet.py ...`` applies the patch and runs the script.  For objects that
already exist, ``install(lit_module)`` swaps them in place.  Nothing here computes on the CPU: the B200 modules raise
when they are called without the CUDA library / device.
"""
import importlib
import sys
import types

_APPLIED = {}


def _import_or_none(name):
    try:
        return importlib.import_module(name)
    except Exception:                                        # missing third-party dependency of an unrelated sub-package
        return None


def _bind_unet(B200Unet):
    bound = []
    smp = _import_or_none("src.segmentation_models")
    if smp is None:
        # stand-in package: only what src/models/encoder_decoder.py:22 needs
        src_pkg = _import_or_none("src")
        if src_pkg is None:
            src_pkg = types.ModuleType("src")
            src_pkg.__path__ = []
            sys.modules["src"] = src_pkg
        smp = types.ModuleType("src.segmentation_models")
        smp.__doc__ = "hallucidet_b200 stand-in for the reference's segmentation_models package (Unet only)"
        smp.__path__ = []
        smp.__b200_stand_in__ = True
        sys.modules["src.segmentation_models"] = smp
        setattr(src_pkg, "segmentation_models", smp)
    for modname in ("src.segmentation_models", "src.segmentation_models.decoders.unet",
                    "src.segmentation_models.decoders.unet.model", "segmentation_models",
                    "segmentation_models.decoders.unet", "segmentation_models.decoders.unet.model"):
        mod = sys.modules.get(modname)
        if mod is not None:
            if modname not in _APPLIED:
                _APPLIED[modname] = getattr(mod, "Unet", None)
            mod.Unet = B200Unet
            bound.append(modname)
    return bound


def _bind_detector(freeze_detector, fast_tail):
    from . import detection
    from .transform import CustomGeneralizedRCNNTransform
    ref = _import_or_none("src.models.detector")
    if ref is None:
        return []
    bound = ["src.models.detector.CustomGeneralizedRCNNTransform"]
    ref.CustomGeneralizedRCNNTransform = CustomGeneralizedRCNNTransform
    cls = ref.Detector
    if not getattr(cls.__init__, "__b200_wrapped__", False):
        original_init = cls.__init__

        def __init__(self, *args, **kwargs):
            original_init(self, *args, **kwargs)
            if freeze_detector:
                # src/models/detector.py:39-66 has built + re-headed the detector; weights loaded later keep their keys
                self.detector = detection.install_b200_backbone(self.detector)

        __init__.__b200_wrapped__ = True
        __init__.__wrapped__ = original_init
        cls.__init__ = __init__
        bound.append("src.models.detector.Detector.__init__")
    if fast_tail:
        ref.eval_forward_fasterrcnn = detection.eval_forward_fasterrcnn
        ref.eval_forward_retinanet = detection.eval_forward_retinanet
        bound += ["src.models.detector.eval_forward_fasterrcnn", "src.models.detector.eval_forward_retinanet"]
        for modname, fn in (("src.utils.eval_forward_fasterrcnn", "eval_forward_fasterrcnn"),
                            ("src.utils.eval_forward_retinanet", "eval_forward_retinanet")):
            mod = sys.modules.get(modname)
            if mod is not None:
                setattr(mod, fn, getattr(detection, fn))
    return bound


def apply(freeze_detector=True, fast_tail=True):
    """Rebind the reference's model constructors to the B200 modules.  Call with the reference checkout on ``sys.path``
    (its scripts run from the repository root), before the models are constructed.  Returns the list of rebound names."""
    from .unet import Unet as B200Unet
    bound = _bind_unet(B200Unet)
    bound += _bind_detector(freeze_detector, fast_tail)
    return bound


def install(lit_module, freeze_detector=True):
    """Swap the models of an already constructed ``EncoderDecoderLit``-like object (attributes ``encoder_decoder`` and
    ``detector``, train_hallucidet.py:80-115) in place: the U-Net is rebuilt as the B200 module from its state dict, the
    detector is frozen and gets the dgrad-only backbone."""
    import torch
    from . import detection
    from .unet import Unet as B200Unet
    unet = getattr(lit_module, "encoder_decoder", None)
    if unet is not None and not isinstance(unet, B200Unet):
        name = getattr(unet, "name", "u-resnet34").split("u-", 1)[-1]
        head = unet.segmentation_head
        new = B200Unet(name, encoder_depth=5, encoder_weights=None, decoder_attention_type=None, in_channels=3,
                       classes=head[0].out_channels)
        new.segmentation_head[-1] = head[-1]
        new.load_state_dict(unet.state_dict(), strict=True)
        new.to(next(unet.parameters()).device)
        new.train(unet.training)
        lit_module.encoder_decoder = new
    det = getattr(lit_module, "detector", None)
    if det is not None and freeze_detector and isinstance(det, torch.nn.Module):
        lit_module.detector = detection.install_b200_backbone(det)
    return lit_module
