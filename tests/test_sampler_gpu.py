"""hd_sample_balanced (hallucidet_b200/csrc/sampler.cu) against torchvision's BalancedPositiveNegativeSampler: the SAME draws
from the same CUDA generator state (torch.randperm's Philox stream restated on the device), and the generator left at the
same offset -- for short rows (every key is a candidate), long rows (key threshold; 64-bit keys above 30083 elements),
tiny rows (islands of equal keys are frequent), empty classes and rows with fewer candidates than the batch size."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from hallucidet_b200 import ops
    return ops


def _reference(labels, bs, frac):
    from torchvision.models.detection._utils import BalancedPositiveNegativeSampler
    pos, neg = BalancedPositiveNegativeSampler(bs, frac)([row for row in labels])
    out = torch.zeros(labels.shape, dtype=torch.uint8, device=labels.device)
    for b, (p, n) in enumerate(zip(pos, neg)):
        out[b][p.bool()] = 1
        out[b][n.bool()] = 2
    return out


def _labels(B, N, p_pos, p_ign, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(B, N, generator=g)
    lab = torch.zeros(B, N)
    lab[u < p_pos] = 1
    lab[(u >= p_pos) & (u < p_pos + p_ign)] = -1
    if dtype == torch.int64:
        lab = lab.to(torch.int64)
        lab[lab == 1] = torch.randint(1, 5, (int((lab == 1).sum()),), generator=g)      # class ids >= 1
    return lab.cuda()


CASES = [
    # B, N, p_pos, p_ign, batch_size_per_image, positive_fraction, dtype
    (8, 2008, 0.05, 0.10, 512, 0.25, torch.int64),        # RoI heads: 2000 proposals + ground truth
    (8, 2008, 0.001, 0.0, 512, 0.25, torch.int64),
    (8, 102300, 0.0005, 0.01, 256, 0.5, torch.float32),   # RPN at 640 x 640: ~100 k negatives -> 64-bit keys + threshold
    (3, 40000, 0.2, 0.1, 256, 0.5, torch.float32),        # 32-bit keys (n <= 30083) with a threshold
    (2, 31000, 0.0, 0.0, 256, 0.5, torch.float32),        # just above the 32 / 64-bit key switch, no positives
    (5, 700, 0.5, 0.3, 512, 0.25, torch.int64),           # fewer candidates than the batch size
    (4, 37, 0.4, 0.2, 16, 0.5, torch.float32),            # tiny rows: islands of equal keys
    (6, 9, 0.5, 0.0, 8, 0.5, torch.int64),
    (2, 5, 1.0, 0.0, 4, 0.5, torch.float32),              # no negatives
    (2, 300, 0.0, 1.0, 64, 0.5, torch.float32),           # everything ignored: no draw at all
]


@pytest.mark.parametrize("B,N,p_pos,p_ign,bs,frac,dtype", CASES)
def test_sample_balanced_matches_torchvision(B, N, p_pos, p_ign, bs, frac, dtype):
    ops = _ops()
    gen = ops.DeviceRng.get(torch.device("cuda", torch.cuda.current_device())).generator()
    for seed in range(6 if N > 1000 else 40):
        labels = _labels(B, N, p_pos, p_ign, 100 + seed, dtype)
        torch.manual_seed(1234 + seed)
        torch.rand(3 + seed, device="cuda")                       # a non-zero starting offset
        start = gen.get_offset()
        ref = _reference(labels, bs, frac)
        ref_offset = gen.get_offset()
        gen.set_offset(start)
        got, counts = ops.sample_balanced(labels, bs, frac)
        ops.DeviceRng.get(labels.device).sync_host()
        assert gen.get_offset() == ref_offset, f"seed {seed}: generator offset {gen.get_offset()} != {ref_offset}"
        assert torch.equal(got, ref), f"seed {seed}: {int((got != ref).sum())} of {got.numel()} entries differ"
        c = counts.cpu()
        assert torch.equal(c[:, 0], (labels >= 1).sum(1).cpu().int()) and torch.equal(c[:, 1], (labels == 0).sum(1).cpu().int())
        assert torch.equal(c[:, 2], (got == 1).sum(1).cpu().int()) and torch.equal(c[:, 3], (got == 2).sum(1).cpu().int())


def test_sample_balanced_chains_across_calls():
    """Two draws in a row without a host sync in between continue the Philox stream on the device (RPN sampler, then RoI
    sampler in one train step) exactly as two torchvision sampler calls on the generator do."""
    ops = _ops()
    gen = ops.DeviceRng.get(torch.device("cuda", torch.cuda.current_device())).generator()
    a = _labels(4, 5000, 0.01, 0.05, 11, torch.float32)
    b = _labels(4, 1200, 0.1, 0.05, 12, torch.int64)
    torch.manual_seed(77)
    ref_a, ref_b = _reference(a, 256, 0.5), _reference(b, 512, 0.25)
    ref_offset = gen.get_offset()
    torch.manual_seed(77)
    got_a, _ = ops.sample_balanced(a, 256, 0.5)
    got_b, _ = ops.sample_balanced(b, 512, 0.25)
    ops.DeviceRng.get(a.device).sync_host()
    assert torch.equal(got_a, ref_a) and torch.equal(got_b, ref_b) and gen.get_offset() == ref_offset


# ---- the sync-free training tail built on the device-side draw (hallucidet_b200.detection STATIC_TAIL) ----------------------

def _detector_setup():
    from oracle import detector as odet
    from torchvision.models.detection.image_list import ImageList
    det = odet.build_detector("fasterrcnn", seed=1).cuda()
    g = torch.Generator().manual_seed(0)
    B = 3
    x = torch.rand(B, 3, 128, 160, generator=g).cuda()
    with torch.no_grad():
        feats = list(det.backbone(x).values())
    anchors = det.rpn.anchor_generator(ImageList(x, [(128, 160)] * B), feats)

    def mk(n):
        xy, wh = torch.rand(n, 2, generator=g) * 80, torch.rand(n, 2, generator=g) * 60 + 8
        return {"boxes": torch.cat([xy, xy + wh], 1).cuda(), "labels": torch.ones(n, dtype=torch.int64).cuda()}
    targets = [mk(3), {"boxes": torch.zeros(0, 4).cuda(), "labels": torch.zeros(0, dtype=torch.int64).cuda()}, mk(5)]
    return det, g, anchors, targets


def test_static_rpn_loss_matches_torchvision():
    """rpn_compute_loss_static (fixed-size row list + validity weights, device-side draw) against RegionProposalNetwork.compute_loss
    on the same generator state: same anchors drawn, losses equal up to the order of the fp32 sums, generator left identical."""
    from hallucidet_b200 import detection as D, ops
    det, g, anchors, targets = _detector_setup()
    B, A = len(anchors), anchors[0].shape[0]
    la, ma = det.rpn.assign_targets_to_anchors(anchors, targets)
    lb, mb = D.assign_targets_to_anchors_batched(det.rpn, anchors, targets)
    obj = torch.randn(B * A, 1, generator=g).cuda().requires_grad_()
    deltas = torch.randn(B * A, 4, generator=g).cuda().requires_grad_()
    rt = det.rpn.box_coder.encode(ma, anchors)
    torch.manual_seed(5)
    l1 = det.rpn.compute_loss(obj, deltas, la, rt)
    g1 = torch.autograd.grad(l1[0] + l1[1], (obj, deltas))
    after1 = torch.rand(1, device="cuda")
    rtb = det.rpn.box_coder.encode_single(mb.reshape(-1, 4), torch.cat(anchors, 0))
    torch.manual_seed(5)
    s = det.rpn.fg_bg_sampler
    sampled, counts = ops.sample_balanced(lb.contiguous(), s.batch_size_per_image, s.positive_fraction)
    l2 = D.rpn_compute_loss_static(det.rpn, obj, deltas, lb, rtb, sampled, counts)
    g2 = torch.autograd.grad(l2[0] + l2[1], (obj, deltas))
    ops.DeviceRng.get(obj.device).sync_host()
    after2 = torch.rand(1, device="cuda")
    assert torch.equal(after1, after2)
    for a, b in zip(l1, l2):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (float(a), float(b))
    for a, b in zip(g1, g2):
        assert torch.equal(a != 0, b != 0) and torch.allclose(a, b, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("sizes", [(50, 37, 64), (900, 1000, 700), (2000, 2000, 1500)])
def test_static_roi_samples_and_loss_match_torchvision(sizes):
    """select_training_samples_static + fastrcnn_loss_masked on padded fixed-shape proposals against torchvision's
    select_training_samples + fastrcnn_loss on the ragged lists: identical sampled proposals / labels / regression targets
    (also when an image has fewer candidates than the 512-row budget: padding rows), losses equal up to summation order."""
    from torchvision.models.detection.roi_heads import fastrcnn_loss
    from hallucidet_b200 import detection as D, ops
    det, g, anchors, targets = _detector_setup()
    rh = det.roi_heads
    T = 2000
    props = [torch.cat([torch.rand(n, 2, generator=g) * 100, torch.rand(n, 2, generator=g) * 60 + 100], 1).cuda() for n in sizes]
    torch.manual_seed(9)
    r_props, r_matched, r_labels, r_reg = rh.select_training_samples([p.clone() for p in props], targets)
    after1 = torch.rand(1, device="cuda")
    padded = torch.zeros(len(props), T, 4, device="cuda")
    for b, p in enumerate(props):
        padded[b, :p.shape[0]] = p
    sp = D._StaticProposals(padded, torch.tensor(sizes, device="cuda"))
    torch.manual_seed(9)
    smp = D.select_training_samples_static(rh, sp, targets)
    ops.DeviceRng.get(padded.device).sync_host()
    after2 = torch.rand(1, device="cuda")
    assert torch.equal(after1, after2)
    n = sum(x.shape[0] for x in r_props)
    assert int(smp.n_drawn) == n and smp.per_image.tolist() == [x.shape[0] for x in r_props]
    assert bool(smp.valid[:n].all()) and not bool(smp.valid[n:].any())
    assert torch.equal(smp.proposals[:n], torch.cat(r_props)) and torch.equal(smp.labels[:n], torch.cat(r_labels))
    assert torch.equal(smp.regression_targets[:n], torch.cat(r_reg)) and torch.equal(smp.matched_idxs[:n], torch.cat(r_matched))
    assert bool((smp.labels[n:] == -100).all())
    img_ref = torch.cat([torch.full((x.shape[0],), b, device="cuda") for b, x in enumerate(r_props)])
    assert torch.equal(smp.image_of[:n], img_ref)
    S = smp.labels.shape[0]
    logits = torch.randn(S, 2, generator=g).cuda().requires_grad_()
    reg = torch.randn(S, 8, generator=g).cuda().requires_grad_()
    l1 = fastrcnn_loss(logits[:n], reg[:n], r_labels, r_reg)
    l2 = D.fastrcnn_loss_masked(logits, reg, smp)
    for a, b in zip(l1, l2):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (float(a), float(b))
    g1 = torch.autograd.grad(l1[0] + l1[1], (logits, reg))
    g2 = torch.autograd.grad(l2[0] + l2[1], (logits, reg))
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-9) and bool((b[n:] == 0).all())


def test_static_tail_matches_host_synchronised_tail():
    """eval_forward_fasterrcnn with STATIC_TAIL on and off, same generator state: same detections, losses equal up to the
    order of fp32 sums -- the sync-free tail draws the same anchors and proposals as the tail that reads the counts on the host."""
    from hallucidet_b200 import detection as D
    from hallucidet_b200.train import HalluciDetTrainer
    from hallucidet_b200.synthetic import synthetic_batch
    dev = torch.device("cuda", 0)
    tr = HalluciDetTrainer(detector_name="fasterrcnn", size=320, seed=3, device=dev)
    ir, rgb, targets = synthetic_batch(4, 256, 320, seed=5, device=dev)
    tr.encoder_decoder.eval()
    with torch.no_grad():
        hal = tr.encoder_decoder(ir.repeat(1, 3, 1, 1)).float()
    outs = {}
    for mode in (False, True):
        D.STATIC_TAIL = mode
        try:
            torch.manual_seed(21)
            x = hal.clone().requires_grad_()
            losses, dets = D.eval_forward_fasterrcnn(tr.detector.model if hasattr(tr.detector, "model") else tr.detector, list(x), targets)
            total = sum(losses.values())
            (gx,) = torch.autograd.grad(total, x)
            outs[mode] = (losses, dets, gx, torch.rand(1, device=dev))
        finally:
            D.STATIC_TAIL = True
    (l0, d0, g0, a0), (l1, d1, g1, a1) = outs[False], outs[True]
    assert torch.equal(a0, a1)
    for k in l0:
        assert torch.allclose(l0[k], l1[k], rtol=1e-5, atol=1e-7), (k, float(l0[k]), float(l1[k]))
    assert len(d0) == len(d1)
    for a, b in zip(d0, d1):
        assert torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["labels"], b["labels"]) and torch.equal(a["scores"], b["scores"])
    rel = float((g0 - g1).norm() / g0.norm())
    assert rel < 2e-2, rel        # bf16 backbone backward: run-to-run gradient noise floor is ~1e-2


@pytest.mark.parametrize("sizes", [(50, 37, 64), (2000, 1800, 1500)])
def test_fused_roi_targets_match_operator_chain(sizes):
    """csrc/roi_targets.cu (hd_roi_match_labels, hd_roi_gather_samples) against the PyTorch operator chain it replaces, bit for
    bit: labels / matched rows of every candidate, then proposals, labels, regression targets, image index, RoIs and pyramid
    levels (torchvision LevelMapper) of the drawn rows."""
    from torchvision.ops import poolers
    from hallucidet_b200 import detection as D, ops
    det, g, anchors, targets = _detector_setup()
    rh = det.roi_heads
    T = 2000
    padded = torch.zeros(len(sizes), T, 4, device="cuda")
    for b, n in enumerate(sizes):
        xy = torch.rand(n, 2, generator=g) * 100
        wh = torch.rand(n, 2, generator=g) ** 3 * 400 + 1                       # sizes from a pixel to several hundred: all pyramid levels
        padded[b, :n] = torch.cat([xy, xy + wh], 1).cuda()
    sp = D._StaticProposals(padded, torch.tensor(sizes, device="cuda"))
    lm = poolers.LevelMapper(2, 5)
    out = {}
    for fused in (False, True):
        D.FUSED_ROI_TARGETS = fused
        try:
            torch.manual_seed(9)
            out[fused] = D.select_training_samples_static(rh, sp, targets, level_mapper=lm)
        finally:
            D.FUSED_ROI_TARGETS = True
    a, b = out[False], out[True]
    assert int(a.n_drawn) == int(b.n_drawn) and torch.equal(a.per_image, b.per_image) and torch.equal(a.valid, b.valid)
    for f in ("proposals", "image_of", "labels", "regression_targets", "matched_idxs"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and torch.equal(x, y), f
    n = int(a.n_drawn)
    assert b.rois is not None and torch.equal(b.rois[:, 0], a.image_of.float()) and torch.equal(b.rois[:, 1:], a.proposals)
    assert torch.equal(b.levels, lm([a.proposals]))
    assert len(torch.unique(b.levels[:n])) >= 3


def test_sample_balanced_sees_a_reseed_without_host_sync():
    """Re-seeding torch's generator to the very state the device chain started from (no sync_host() in between) restarts the
    draw: the pending marker offset makes the re-seed visible."""
    ops = _ops()
    lab = _labels(4, 3000, 0.02, 0.05, 31, torch.float32)
    torch.manual_seed(5)
    a1, _ = ops.sample_balanced(lab, 256, 0.5)
    torch.manual_seed(5)
    a2, _ = ops.sample_balanced(lab, 256, 0.5)
    a3, _ = ops.sample_balanced(lab, 256, 0.5)                    # no re-seed: the chain continues
    ops.DeviceRng.get(lab.device).sync_host()
    torch.manual_seed(5)
    ref1, ref2 = _reference(lab, 256, 0.5), _reference(lab, 256, 0.5)
    assert torch.equal(a1, a2) and torch.equal(a1, ref1) and torch.equal(a3, ref2) and not torch.equal(a1, a3)


def test_rpn_concat_preds_matches_torchvision_and_autograd():
    """hd_rpn_concat_preds (both directions) against torchvision's concat_box_prediction_layers on NCHW views of the same
    channels-last predictor maps, forward values and the gradients autograd sends back."""
    from torchvision.models.detection.rpn import concat_box_prediction_layers
    from hallucidet_b200 import heads
    g = torch.Generator().manual_seed(4)
    a, B = 3, 2
    preds = [torch.randn(B, h, w, 16, generator=g).cuda().requires_grad_() for h, w in ((20, 24), (10, 12), (5, 6), (3, 3), (2, 2))]
    obj, deltas = heads._ConcatRPNPreds.apply(a, None, None, *preds)
    ref_preds = [p.detach().clone().requires_grad_() for p in preds]
    nchw = [p.permute(0, 3, 1, 2) for p in ref_preds]
    ro, rd = concat_box_prediction_layers([x[:, :a] for x in nchw], [x[:, a:5 * a] for x in nchw])
    assert torch.equal(obj, ro) and torch.equal(deltas, rd)
    wo, wd = torch.randn(obj.shape, generator=g).cuda(), torch.randn(deltas.shape, generator=g).cuda()
    ((obj * wo).sum() + (deltas * wd).sum()).backward()
    ((ro * wo).sum() + (rd * wd).sum()).backward()
    for p, q in zip(preds, ref_preds):
        assert torch.equal(p.grad, q.grad)


def test_roi_align_ml_fwd_bf16_equals_fp32_on_widened_maps():
    """hd_roi_align_ml_fwd_bf16 reads the bf16 pyramid; on fp32 maps that are widened copies of it the fp32 entry point gives the
    same bits (same arithmetic, half the traffic)."""
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    B, C = 2, 256
    sizes = ((40, 48), (20, 24), (10, 12), (5, 6))
    scales = (0.25, 0.125, 0.0625, 0.03125)
    maps16 = [torch.randn(B, h, w, C, generator=g).to(torch.bfloat16).cuda() for h, w in sizes]
    maps32 = [m.float() for m in maps16]
    K = 300
    xy = torch.rand(K, 2, generator=g) * 120
    wh = torch.rand(K, 2, generator=g) ** 2 * 150 + 2
    rois = torch.cat([torch.randint(0, B, (K, 1), generator=g).float(), xy, xy + wh], 1).cuda()
    levels = torch.randint(0, 4, (K,), generator=g).cuda()
    a = ops.roi_align_ml_fwd(maps32, scales, rois, levels, (7, 7), 2)
    b = ops.roi_align_ml_fwd(maps16, scales, rois, levels, (7, 7), 2)
    assert torch.equal(a, b)


def test_fused_proposal_decode_matches_decode_all_then_gather():
    """filter_proposals_static_fused (top-k first, then ONE launch decoding / clipping / testing only the selected candidates)
    against filter_proposals_static on BoxCoder.decode of every anchor: identical proposals and counts, bit for bit."""
    from torchvision.models.detection.image_list import ImageList
    from hallucidet_b200 import detection as D
    det, g, anchors, targets = _detector_setup()
    B = len(anchors)
    x = torch.rand(B, 3, 128, 160, generator=g).cuda()
    with torch.no_grad():
        feats = list(det.backbone(x).values())
        obj, deltas = det.rpn.head(feats)
    napl = [o[0].shape[0] * o[0].shape[1] * o[0].shape[2] for o in obj]
    from torchvision.models.detection.rpn import concat_box_prediction_layers
    o2, d2 = concat_box_prediction_layers(obj, deltas)
    d2 = d2 * 3.0                                                   # wider spread of box sizes: clipping / min-size tests bite
    sizes = [(128, 160)] * B
    props = D._decode(det.rpn.box_coder, d2, anchors).view(B, -1, 4)
    a = D.filter_proposals_static(det.rpn, props, o2, sizes, napl)
    b = D.filter_proposals_static_fused(det.rpn, d2, anchors, o2, sizes, napl)
    assert torch.equal(a.counts, b.counts) and torch.equal(a.boxes, b.boxes)
    assert int(a.counts.min()) > 50


def test_fused_rpn_targets_match_operator_chain():
    """hd_rpn_assign_targets against assign_targets_to_anchors (torchvision, per image) + BoxCoder.encode: labels of every anchor
    and the regression targets of every anchor that has a matched box, bit for bit (one image without ground truth)."""
    from hallucidet_b200 import detection as D
    det, g, anchors, targets = _detector_setup()
    la, ma = det.rpn.assign_targets_to_anchors(anchors, targets)
    rt = det.rpn.box_coder.encode(ma, anchors)
    labels, reg = D.rpn_targets_static(det.rpn, anchors, targets)
    A = anchors[0].shape[0]
    assert torch.equal(labels, torch.stack(la))
    assert sum(int((x == 1).sum()) for x in la) > 0 and sum(int((x == -1).sum()) for x in la) > 0
    reg = reg.view(len(anchors), A, 4)
    for b in range(len(anchors)):
        if targets[b]["boxes"].numel() == 0:
            continue
        assert torch.equal(reg[b], rt[b])


@pytest.mark.parametrize("B,N,p_pos,p_ign,bs,frac,dtype", [(3, 700, 0.05, 0.1, 64, 0.5, torch.float32), (2, 45000, 0.001, 0.0, 256, 0.5, torch.float32),
                                                           (4, 50, 0.4, 0.2, 16, 0.25, torch.int64)])
def test_sample_balanced_matches_the_oracle(B, N, p_pos, p_ign, bs, frac, dtype):
    """hd_sample_balanced against oracle/randperm_cuda.py (the numpy restatement of torch.randperm's CUDA algorithm, pinned by
    golden vectors of torch itself): same draws and same generator offset for an explicit (seed, offset) state."""
    from oracle import randperm_cuda as R
    ops = _ops()
    gen = ops.DeviceRng.get(torch.device("cuda", torch.cuda.current_device())).generator()
    labels = _labels(B, N, p_pos, p_ign, 77, dtype)
    seed, offset = 20240607, 1024
    gen.manual_seed(seed)
    gen.set_offset(offset)
    got, _ = ops.sample_balanced(labels, bs, frac)
    ops.DeviceRng.get(labels.device).sync_host()
    want, after = R.balanced_sample(labels.cpu().numpy(), bs, frac, seed, offset)
    assert gen.get_offset() == after
    assert torch.equal(got.cpu(), torch.from_numpy(want))
