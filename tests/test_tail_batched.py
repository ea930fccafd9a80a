"""Whole-batch target assignment / sampling (hallucidet_b200.detection) against torchvision's per-image loops: same labels,
same matched boxes, same random samples (generator consumed identically), same losses.  Runs on the CPU (host logic) and,
marked gpu, on CUDA, where the CUDA generator and the CUDA arg-max tie rules are what matters."""
import pytest
import torch

from oracle import detector as odet


def _setup(device):
    from torchvision.models.detection.image_list import ImageList
    det = odet.build_detector("fasterrcnn", seed=1).to(device)
    g = torch.Generator().manual_seed(0)
    B = 3
    x = torch.rand(B, 3, 128, 160, generator=g).to(device)
    with torch.no_grad():
        feats = list(det.backbone(x).values())
    il = ImageList(x, [(128, 160)] * B)
    anchors = det.rpn.anchor_generator(il, feats)

    def mk(n):
        xy, wh = torch.rand(n, 2, generator=g) * 80, torch.rand(n, 2, generator=g) * 60 + 8
        return {"boxes": torch.cat([xy, xy + wh], 1).to(device), "labels": torch.ones(n, dtype=torch.int64, device=device)}
    targets = [mk(3), {"boxes": torch.zeros(0, 4, device=device), "labels": torch.zeros(0, dtype=torch.int64, device=device)}, mk(5)]
    return det, g, anchors, targets


def _check(device):
    from hallucidet_b200 import detection as D
    det, g, anchors, targets = _setup(device)
    B, A = len(anchors), anchors[0].shape[0]
    la, ma = det.rpn.assign_targets_to_anchors(anchors, targets)
    lb, mb = D.assign_targets_to_anchors_batched(det.rpn, anchors, targets)
    assert all(torch.equal(a, b) for a, b in zip(la, lb)) and all(torch.equal(a, b) for a, b in zip(ma, mb))
    assert sum(int((l == 1).sum()) for l in la) > 0 and sum(int((l == -1).sum()) for l in la) > 0
    obj = torch.randn(B * A, 1, generator=g).to(device)
    deltas = torch.randn(B * A, 4, generator=g).to(device)
    rt = det.rpn.box_coder.encode(ma, anchors)
    torch.manual_seed(5)
    l1 = det.rpn.compute_loss(obj, deltas, la, rt)
    after1 = torch.rand(1, device=device)
    rtb = det.rpn.box_coder.encode_single(mb.reshape(-1, 4), torch.cat(anchors, 0))
    torch.manual_seed(5)
    l2 = D.rpn_compute_loss_batched(det.rpn, obj, deltas, lb, rtb)
    after2 = torch.rand(1, device=device)
    assert torch.equal(l1[0], l2[0]) and torch.equal(l1[1], l2[1])
    assert torch.equal(after1, after2)                   # the generator is left in the same state
    # RoI sampling: ragged proposal counts, one image without ground truth, more candidates than the 512-sample budget
    for sizes in ((50, 37, 64), (900, 1000, 700)):
        props = [torch.cat([torch.rand(n, 2, generator=g) * 100, torch.rand(n, 2, generator=g) * 60 + 100], 1).to(device) for n in sizes]
        torch.manual_seed(9)
        r1 = det.roi_heads.select_training_samples([p.clone() for p in props], targets)
        torch.manual_seed(9)
        r2 = D.select_training_samples_batched(det.roi_heads, [p.clone() for p in props], targets)
        for a, b in zip(r1, r2):
            assert [x.shape for x in a] == [y.shape for y in b]
            assert all(torch.equal(x, y) for x, y in zip(a, b))


def _check_roi_pool_and_loss(device):
    from collections import OrderedDict
    from torchvision.models.detection.roi_heads import fastrcnn_loss
    from hallucidet_b200 import detection as D
    det = odet.build_detector("fasterrcnn", seed=1).to(device)
    g = torch.Generator().manual_seed(0)
    feats = OrderedDict((k, torch.randn(2, 256, s, s, generator=g).to(device)) for k, s in (("0", 64), ("1", 32), ("2", 16), ("3", 8), ("pool", 4)))

    def boxes(n):
        xy, wh = torch.rand(n, 2, generator=g) * 150, torch.rand(n, 2, generator=g) * 100 + 2
        return torch.cat([xy, xy + wh], 1).to(device)
    props = [boxes(200), boxes(137)]
    a = det.roi_heads.box_roi_pool(feats, props, [(256, 256)] * 2)
    b = D.multiscale_roi_align_one_sync(det.roi_heads.box_roi_pool, feats, props, [(256, 256)] * 2)
    assert torch.equal(a, b) and a.shape == (337, 256, 7, 7)
    for fused in ((True, False) if device == "cuda" else ()):   # with gradients: the hd_roi_align_* kernels run (forward bit-identical,
        D.ROI_ALIGN_FUSED_LEVELS = fused                         # backward close), all levels in one launch or level by level
        try:
            fg = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in feats.items())
            fr = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in feats.items())
            w = torch.randn(a.shape, generator=g).to(device)
            c = D.multiscale_roi_align_one_sync(det.roi_heads.box_roi_pool, fg, props, [(256, 256)] * 2)
            r = det.roi_heads.box_roi_pool(fr, props, [(256, 256)] * 2)
            assert torch.equal(c, r)
            (c * w).sum().backward()
            (r * w).sum().backward()
            for k in ("0", "1", "2", "3"):
                assert torch.allclose(fg[k].grad, fr[k].grad, rtol=1e-4, atol=1e-5 * float(fr[k].grad.abs().max()) + 1e-12)
            assert fg["pool"].grad is None
        finally:
            D.ROI_ALIGN_FUSED_LEVELS = True
    cl, br = torch.randn(337, 2, generator=g).to(device), torch.randn(337, 8, generator=g).to(device)
    labels = [(torch.rand(n, generator=g) > 0.7).long().to(device) for n in (200, 137)]
    rt = [torch.randn(n, 4, generator=g).to(device) for n in (200, 137)]
    l1 = fastrcnn_loss(cl, br, labels, rt)
    l2 = D.fastrcnn_loss_static(cl, br, labels, rt, sum(int((l > 0).sum()) for l in labels))
    assert torch.equal(l1[0], l2[0]) and torch.equal(l1[1], l2[1])


def test_roi_pool_and_loss_cpu():
    _check_roi_pool_and_loss("cpu")


@pytest.mark.gpu
def test_roi_pool_and_loss_cuda():
    _check_roi_pool_and_loss("cuda")


def test_batched_targets_and_sampling_cpu():
    _check("cpu")


@pytest.mark.gpu
def test_batched_targets_and_sampling_cuda():
    _check("cuda")


@pytest.mark.gpu
def test_retinanet_tail_cuda():
    """RetinaNet (BASELINE config 4): batched anchor matching + count-known index lists give the same losses as the
    per-image loop, and the hd_nms-based post-processing returns torchvision's detections."""
    from torchvision.models.detection.image_list import ImageList
    from hallucidet_b200 import detection as D
    det = odet.build_detector("retinanet", seed=1).cuda().eval()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, 3, 128, 160, generator=g).cuda()

    def mk(n):
        xy, wh = torch.rand(n, 2, generator=g) * 80, torch.rand(n, 2, generator=g) * 60 + 8
        return {"boxes": torch.cat([xy, xy + wh], 1).cuda(), "labels": torch.ones(n, dtype=torch.int64, device="cuda")}
    targets = [mk(3), {"boxes": torch.zeros(0, 4, device="cuda"), "labels": torch.zeros(0, dtype=torch.int64, device="cuda")}, mk(5)]
    with torch.no_grad():
        feats = list(det.backbone(x).values())
        ho = det.head(feats)
        il = ImageList(x, [(128, 160)] * 3)
        anchors = det.anchor_generator(il, feats)
        a = D.compute_retinanet_loss(targets, ho, anchors, det, batched=False)
        D.WHOLE_BATCH_RETINANET_LOSS = False                      # count-known index lists: same terms in the same order
        try:
            b = D.compute_retinanet_loss(targets, ho, anchors, det, batched=True)
        finally:
            D.WHOLE_BATCH_RETINANET_LOSS = True
        assert torch.equal(a["classification"], b["classification"]) and torch.equal(a["bbox_regression"], b["bbox_regression"])
        c = D.compute_retinanet_loss(targets, ho, anchors, det, batched=True)   # whole-batch, sync-free: other fp32 summation order
        assert torch.allclose(a["classification"], c["classification"], rtol=1e-5) and torch.allclose(a["bbox_regression"], c["bbox_regression"], rtol=1e-5)
        assert float(a["bbox_regression"]) > 0
        napl = [f.size(2) * f.size(3) for f in feats]
        per = ho["cls_logits"].size(1) // sum(napl)
        napl = [n * per for n in napl]
        sho = {k: list(v.split(napl, dim=1)) for k, v in ho.items()}
        sa = [list(t.split(napl)) for t in anchors]
        det.score_thresh = 0.0095                       # random-init scores sit around the 0.01 prior
        d1 = det.postprocess_detections(sho, sa, il.image_sizes)
        d2 = D.retinanet_postprocess_detections(det, sho, sa, il.image_sizes)
        b3, s3, l3 = D._resolve(D.retinanet_postprocess_detections_batched_begin(det, sho, sa, il.image_sizes))[0]
    assert sum(d["boxes"].shape[0] for d in d1) > 0
    for i, (p, q) in enumerate(zip(d1, d2)):
        assert all(torch.equal(p[k], q[k]) for k in ("boxes", "scores", "labels"))
        assert torch.equal(p["boxes"], b3[i]) and torch.equal(p["scores"], s3[i]) and torch.equal(p["labels"], l3[i])


def test_pad_rows_matches_pad_sequence_cpu():
    """detection._pad_rows (one concatenation + index_copy) == row-by-row padding, incl. empty rows and several pieces per image."""
    from hallucidet_b200 import detection as D
    g = torch.Generator().manual_seed(3)
    rows = [torch.randn(n, 4, generator=g) for n in (5, 0, 3, 7)]
    out, present = D._pad_rows(rows, 9, fill=-1.0)
    ref = torch.full((4, 9, 4), -1.0)
    for b, r in enumerate(rows):
        ref[b, :r.shape[0]] = r
    assert torch.equal(out, ref)
    assert torch.equal(present, torch.arange(9)[None, :] < torch.tensor([5, 0, 3, 7])[:, None])
    # all rows full: stack
    full = [torch.randn(6, 4, generator=g) for _ in range(3)]
    out, present = D._pad_rows(full, 6)
    assert torch.equal(out, torch.stack(full)) and bool(present.all())
    # nothing at all
    out, present = D._pad_rows([torch.zeros(0, 4), torch.zeros(0, 4)], 2, fill=0)
    assert out.shape == (2, 2, 4) and not bool(present.any()) and float(out.abs().sum()) == 0.0
    # two pieces per image (proposals followed by ground truth), image-major
    pieces = [torch.randn(n, 4, generator=g) for n in (4, 2, 0, 1, 3, 0)]
    out, present = D._pad_rows(pieces, 7, counts=[6, 1, 3])
    ref = torch.zeros(3, 7, 4)
    ref[0, :6] = torch.cat(pieces[0:2]); ref[1, :1] = torch.cat(pieces[2:4]); ref[2, :3] = torch.cat(pieces[4:6])
    assert torch.equal(out, ref) and present.sum(1).tolist() == [6, 1, 3]
    labels = [torch.arange(n) + 10 for n in (2, 0, 4)]
    out, _ = D._pad_rows(labels, 4)
    assert out.dtype == torch.int64 and out.tolist() == [[10, 11, 0, 0], [0, 0, 0, 0], [10, 11, 12, 13]]


def test_padded_gt_cache_follows_the_targets_cpu():
    from hallucidet_b200 import detection as D
    t = [{"boxes": torch.tensor([[0., 0., 4., 4.], [1., 1., 3., 5.]]), "labels": torch.tensor([1, 1])},
         {"boxes": torch.zeros(0, 4), "labels": torch.zeros(0, dtype=torch.int64)}]
    gt, present, gl = D._padded_gt(t, torch.float32)
    assert gt.shape == (2, 2, 4) and present.tolist() == [[True, True], [False, False]] and gl.tolist() == [[1, 1], [0, 0]]
    assert D._padded_gt(t, torch.float32)[0] is gt                          # same targets: cached
    t[0]["boxes"][0, 2] = 9.0                                                # in-place change bumps the version
    gt2, _, _ = D._padded_gt(t, torch.float32)
    assert gt2 is not gt and float(gt2[0, 0, 2]) == 9.0
    t2 = [{"boxes": x["boxes"].clone(), "labels": x["labels"].clone()} for x in t]
    assert D._padded_gt(t2, torch.float32)[0] is not gt2                     # other tensors: recomputed


def test_fused_head_modules_keep_state_dict_and_cpu_results():
    """install_b200_backbone swaps Conv2dNormActivation(conv, ReLU) of the frozen heads for detection._FusedConvReLU: the
    state_dict keys do not change and, off the GPU, the module computes exactly what it replaced."""
    import copy
    from torchvision.models.detection.rpn import RPNHead
    from torchvision.models.detection.retinanet import RetinaNetClassificationHead
    from hallucidet_b200 import detection as D
    for head in (RPNHead(16, 3), RetinaNetClassificationHead(16, 3, 2)):
        for p in head.parameters():
            p.requires_grad_(False)
        fused = copy.deepcopy(head)
        D._fuse_head_conv_relu(fused)
        assert any(isinstance(m, D._FusedConvReLU) for m in fused.modules())
        assert list(fused.state_dict()) == list(head.state_dict())
        xs = [torch.randn(2, 16, 12, 10), torch.randn(2, 16, 6, 5)]
        a, b = head(xs), fused(xs)
        a = a if isinstance(a, torch.Tensor) else [t for part in a for t in part]
        b = b if isinstance(b, torch.Tensor) else [t for part in b for t in part]
        if isinstance(a, torch.Tensor):
            assert torch.equal(a, b)
        else:
            assert all(torch.equal(x, y) for x, y in zip(a, b))
    rpn = RPNHead(16, 3)
    xs = [torch.randn(1, 16, 8, 8)]
    o1, d1 = rpn(xs)
    o2, d2 = D._rpn_head(rpn, xs)                                            # CPU: falls back to the module
    assert torch.equal(o1[0], o2[0]) and torch.equal(d1[0], d2[0])
