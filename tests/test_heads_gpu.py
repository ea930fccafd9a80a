"""Frozen detector heads on the B200 conv kernels (hallucidet_b200/heads.py) against torchvision's own
head modules in fp32 (TF32 off) on the same bf16-representable pyramid: predictor outputs and the input gradients
(teacher-forced: identical inputs, so only bf16 storage of the hidden maps and accumulation order differ)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _pyramid(sizes, b=2, c=256, seed=0):
    g = torch.Generator().manual_seed(seed)
    bf16 = [(torch.randn(b, h, w, c, generator=g)).to(torch.bfloat16).cuda() for h, w in sizes]
    f32 = [x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True) for x in bf16]
    f32b = [x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True) for x in bf16]
    return bf16, f32, f32b


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _check(mine, ref, what, rtol=1.5e-2):
    err = (mine.float() - ref.float()).abs().max().item()
    lim = rtol * ref.abs().max().item() + 1e-6
    assert err <= lim, f"{what}: max err {err:.4g} > {lim:.4g}"


def _boost(module, gain):
    """torchvision initialises the head convs with std 0.01: scale them so that rounding errors would be visible, and make
    them bf16-representable: with fp32-only weight bits the hidden maps differ by ~2e-3 relative, which flips the ReLU mask
    of ~0.1 % of the (near-zero) elements -- each flip is a full-magnitude error in the gradient (3 % in L2, measured), the
    same mechanism that makes free-running end-to-end gradients incomparable."""
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.Conv2d):
                m.weight.copy_((m.weight * gain).to(torch.bfloat16).float())
                m.bias.normal_(0, 0.1)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


@pytest.mark.parametrize("sizes", [[(40, 40), (20, 20), (10, 10), (5, 5), (3, 3)], [(75, 75), (38, 38), (19, 19), (10, 10), (5, 5)]])
def test_rpn_head_matches_torchvision(sizes):
    from hallucidet_b200 import heads
    from oracle import detector as odet
    det = odet.build_detector("fasterrcnn", seed=123).cuda()
    head = det.rpn.head
    _boost(head, 4.0)
    bf16, f32, f32b = _pyramid(sizes)
    ref_l, ref_b = head(f32)
    assert heads.rpn_head_tower(head) is not None
    my_l, my_b = heads.rpn_head_forward(head, f32b, bf16)
    gen = torch.Generator().manual_seed(5)
    loss_r = loss_m = 0.0
    for rl, rb, ml, mb in zip(ref_l, ref_b, my_l, my_b):
        assert ml.shape == rl.shape and mb.shape == rb.shape
        _check(ml, rl.detach(), "objectness")
        _check(mb, rb.detach(), "bbox deltas")
        w1 = torch.randn(rl.shape, generator=gen).cuda()
        w2 = torch.randn(rb.shape, generator=gen).cuda()
        loss_r = loss_r + (rl * w1).sum() + (rb * w2).sum()
        loss_m = loss_m + (ml * w1).sum() + (mb * w2).sum()
    loss_r.backward()
    loss_m.backward()
    for a, b in zip(f32b, f32):
        assert a.grad is not None and a.grad.shape == b.grad.shape
        assert _rel(a.grad, b.grad) <= 1.5e-2, _rel(a.grad, b.grad)
        assert _cos(a.grad, b.grad) >= 0.9998


def test_retinanet_head_matches_torchvision():
    from hallucidet_b200 import heads
    from oracle import detector as odet
    det = odet.build_detector("retinanet", seed=123).cuda()
    head = det.head
    _boost(head, 3.0)
    sizes = [(40, 40), (20, 20), (10, 10), (5, 5), (3, 3)]
    bf16, f32, f32b = _pyramid(sizes)
    ref = head(f32)
    assert heads.retinanet_head_towers(head) is not None
    mine = heads.retinanet_head_forward(head, f32b, bf16)
    gen = torch.Generator().manual_seed(5)
    loss_r = loss_m = 0.0
    for key in ("cls_logits", "bbox_regression"):
        assert mine[key].shape == ref[key].shape
        _check(mine[key], ref[key].detach(), key, rtol=2.5e-2)
        w = torch.randn(ref[key].shape, generator=gen).cuda()
        loss_r = loss_r + (ref[key] * w).sum()
        loss_m = loss_m + (mine[key] * w).sum()
    loss_r.backward()
    loss_m.backward()
    for a, b in zip(f32b, f32):
        # 4-conv towers: every hidden map is stored in bf16 (2^-9 relative), so each further conv sees ~3e-3 different inputs and
        # ~0.1 % of its near-zero outputs change sign -> ReLU-mask flips, each a full-magnitude gradient error (7 % in L2 measured)
        assert _rel(a.grad, b.grad) <= 0.12, _rel(a.grad, b.grad)
        assert _cos(a.grad, b.grad) >= 0.995


@pytest.mark.parametrize("name", ["fasterrcnn", "retinanet"])
def test_train_step_with_b200_heads_vs_cudnn_heads(name):
    """The assembled step with the heads on the B200 kernels vs the same step with torchvision's head modules (cuDNN fp32):
    total loss within 1 %, U-Net gradient direction preserved."""
    from hallucidet_b200 import detection
    from hallucidet_b200.synthetic import synthetic_batch
    from hallucidet_b200.train import HalluciDetTrainer
    from oracle import detector as odet
    dev = torch.device("cuda", 0)
    ir, rgb, targets = synthetic_batch(2, 128, 160, seed=123, device=dev)
    det_cpu = odet.build_detector(name, seed=123)
    odet.randomize_bn_stats(det_cpu, seed=7)
    res = {}
    for flag in (True, False):
        detection.B200_HEADS = flag
        try:
            tr = HalluciDetTrainer(detector_name=name, size=160, seed=123, device=dev, detector_state=det_cpu.state_dict())
            tr.encoder_decoder.train()
            out = tr.forward_step(rgb, targets, ir, targets, det_seed=7)
            out["total"].backward()
            torch.cuda.synchronize()
            res[flag] = (float(out["total"].detach()), torch.cat([p.grad.flatten() for p in tr.encoder_decoder.parameters()]).clone())
        finally:
            detection.B200_HEADS = True
    rel = abs(res[True][0] - res[False][0]) / abs(res[False][0])
    c = _cos(res[True][1], res[False][1])
    print(f"\n[{name}] loss with B200 heads {res[True][0]:.6f} vs cuDNN heads {res[False][0]:.6f} (rel {rel:.2e}); grad cosine {c:.4f}")
    assert rel <= 1e-2
    assert c >= 0.9            # (run-to-run noise of the free-running gradient alone is ~1e-2 relative, see test_modules_gpu)


def test_rpn_head_cuda_graph_program_matches_eager():
    """The head tower as two CUDA graphs on static buffers (heads._TowerProgram): warm-up, capture and replay all give the
    eager launches' results, for changing inputs in the same (static) pyramid buffers."""
    from hallucidet_b200 import heads
    from oracle import detector as odet
    det = odet.build_detector("fasterrcnn", seed=123).cuda()
    head = det.rpn.head
    _boost(head, 4.0)
    sizes = [(40, 40), (20, 20), (10, 10), (5, 5), (3, 3)]
    bf16, _, _ = _pyramid(sizes)
    gen = torch.Generator().manual_seed(9)
    try:
        for it in range(4):
            for x in bf16:                                          # new values, same buffers (as the backbone engine's pyramid)
                x.copy_(torch.randn(x.shape, generator=gen).to(torch.bfloat16))
            res = {}
            for graph in (False, True):
                heads.USE_CUDA_GRAPH = graph
                f32 = [x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True) for x in bf16]
                lo, bb = heads.rpn_head_forward(head, f32, bf16)
                loss = sum((a * (i + 1)).sum() + (b * 0.5).sum() for i, (a, b) in enumerate(zip(lo, bb)))
                loss.backward()
                torch.cuda.synchronize()
                res[graph] = ([t.detach().clone() for t in lo + bb], [f.grad.clone() for f in f32])
            for a, b in zip(res[False][0] + res[False][1], res[True][0] + res[True][1]):
                assert torch.equal(a, b), it
        # the program must really have been selected and captured (its selection once tested the grad mode inside
        # Function.forward, where it is always off, and silently ran the eager launches instead)
        tower = heads.rpn_head_tower(head)
        prog = tower.programs.get(tuple(x.data_ptr() for x in bf16))
        assert prog is not None and isinstance(prog.graphs.get("fwd"), torch.cuda.CUDAGraph)
        assert isinstance(prog.graphs.get("bwd"), torch.cuda.CUDAGraph)
    finally:
        heads.USE_CUDA_GRAPH = False


def test_retinanet_deferred_detections_equal_inline():
    """Train-step mode post-processes RetinaNet's detections on a side stream after the backward pass has been enqueued
    (detection.DeferredCall): same detections and losses as the in-line path."""
    from hallucidet_b200 import detection
    from hallucidet_b200.synthetic import synthetic_batch
    from hallucidet_b200.train import HalluciDetTrainer
    from oracle import detector as odet
    dev = torch.device("cuda", 0)
    ir, rgb, targets = synthetic_batch(2, 128, 160, seed=123, device=dev)
    det_cpu = odet.build_detector("retinanet", seed=123)
    odet.randomize_bn_stats(det_cpu, seed=7)
    tr = HalluciDetTrainer(detector_name="retinanet", size=160, seed=123, device=dev, detector_state=det_cpu.state_dict())
    tr.encoder_decoder.eval()
    with torch.no_grad():
        hal = tr.encoder_decoder(ir.expand(-1, 3, -1, -1))
    res = {}
    for defer in (False, True):
        detection.DEFER_DETECTIONS = defer
        try:
            hal_in = hal.clone().requires_grad_(True)
            losses, det = detection.Detector.calculate_loss(tr.detector, hal_in, targets, model_name="retinanet")
            (losses["classification"] + losses["bbox_regression"]).backward()
            if defer:
                assert isinstance(det, detection.DeferredDetections)
                det = det.resolve()
            torch.cuda.synchronize()
            res[defer] = (losses, det)
        finally:
            detection.DEFER_DETECTIONS = False
    for k in ("classification", "bbox_regression"):
        assert torch.equal(res[False][0][k], res[True][0][k])
    assert len(res[False][1]) == len(res[True][1]) == 2
    for a, b in zip(res[False][1], res[True][1]):
        for k in ("boxes", "scores", "labels"):
            assert torch.equal(a[k], b[k]), k


def test_retinanet_batched_postprocess_equals_torchvision():
    """retinanet_postprocess_detections_batched (whole batch per level, one host sync) against torchvision's own
    ``RetinaNet.postprocess_detections`` on the same head outputs: identical boxes / scores / labels."""
    from hallucidet_b200 import detection
    from oracle import detector as odet
    from torchvision.models.detection.image_list import ImageList
    det = odet.build_detector("retinanet", seed=123).cuda()
    det.score_thresh = 0.3
    B, sizes = 3, [(20, 20), (10, 10), (5, 5), (3, 3), (2, 2)]
    g = torch.Generator().manual_seed(3)
    feats = [torch.randn(B, 256, h, w, generator=g).cuda() for h, w in sizes]
    images = ImageList(torch.zeros(B, 3, 160, 160, device="cuda"), [(160, 160)] * B)
    anchors = det.anchor_generator(images, feats)
    A = det.anchor_generator.num_anchors_per_location()[0]
    n_per = [h * w * A for h, w in sizes]
    total = sum(n_per)
    head = {"cls_logits": (torch.randn(B, total, 2, generator=g) * 2).cuda(), "bbox_regression": (torch.randn(B, total, 4, generator=g) * 0.3).cuda()}
    split_head = {k: list(v.split(n_per, dim=1)) for k, v in head.items()}
    split_anchors = [list(a.split(n_per)) for a in anchors]
    ref = det.postprocess_detections(split_head, split_anchors, images.image_sizes)
    boxes, scores, labels = detection._resolve(detection.retinanet_postprocess_detections_batched_begin(det, split_head, split_anchors, images.image_sizes))[0]
    assert sum(len(r["boxes"]) for r in ref) > 50
    for i, r in enumerate(ref):
        assert torch.equal(boxes[i], r["boxes"]) and torch.equal(scores[i], r["scores"]) and torch.equal(labels[i], r["labels"]), i


def test_retinanet_whole_batch_loss_matches_per_image_loop():
    """compute_retinanet_loss: the sync-free whole-batch form against the per-image loop (the reference's
    src/utils/eval_forward_retinanet.py:163-244 order of operations): same losses and gradients up to fp32 summation order."""
    from hallucidet_b200 import detection
    from oracle import detector as odet
    from torchvision.models.detection.image_list import ImageList
    det = odet.build_detector("retinanet", seed=123).cuda()
    B, sizes = 3, [(20, 20), (10, 10), (5, 5), (3, 3), (2, 2)]
    g = torch.Generator().manual_seed(4)
    feats = [torch.randn(B, 256, h, w, generator=g).cuda() for h, w in sizes]
    images = ImageList(torch.zeros(B, 3, 160, 160, device="cuda"), [(160, 160)] * B)
    anchors = det.anchor_generator(images, feats)
    total = anchors[0].shape[0]
    targets = [{"boxes": torch.tensor([[10., 12., 60., 90.], [70., 40., 150., 120.]]).cuda(), "labels": torch.tensor([1, 1]).cuda()},
               {"boxes": torch.tensor([[5., 5., 40., 40.]]).cuda(), "labels": torch.tensor([1]).cuda()},
               {"boxes": torch.zeros(0, 4).cuda(), "labels": torch.zeros(0, dtype=torch.int64).cuda()}]
    res = {}
    for flag in (False, True):
        detection.WHOLE_BATCH_RETINANET_LOSS = flag
        try:
            logits = (torch.randn(B, total, 2, generator=torch.Generator().manual_seed(5)) * 2).cuda().requires_grad_(True)
            reg = (torch.randn(B, total, 4, generator=torch.Generator().manual_seed(6)) * 0.3).cuda().requires_grad_(True)
            losses = detection.compute_retinanet_loss(targets, {"cls_logits": logits, "bbox_regression": reg}, anchors, det)
            (losses["classification"] + 2.0 * losses["bbox_regression"]).backward()
            res[flag] = (losses, logits.grad.clone(), reg.grad.clone())
        finally:
            detection.WHOLE_BATCH_RETINANET_LOSS = True
    for k in ("classification", "bbox_regression"):
        assert torch.allclose(res[True][0][k], res[False][0][k], rtol=1e-5, atol=1e-7), k
    assert torch.allclose(res[True][1], res[False][1], rtol=1e-4, atol=1e-8)
    assert torch.allclose(res[True][2], res[False][2], rtol=1e-4, atol=1e-8)
