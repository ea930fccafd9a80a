"""Host-side formulas of the sync-free training tail (hallucidet_b200.detection) on the CPU: the fixed-size row list + validity
weights that replace torchvision's data-dependent index lists must give torchvision's losses.  (The device-side draw and the
fused kernels behind the same functions are covered by tests/test_sampler_gpu.py on the GPU.)"""
import torch

from hallucidet_b200 import detection as D


def _sampled_from_torchvision(labels, bs, frac, seed):
    """sampled [B, N] uint8 (1 positive / 2 negative drawn) + counts [B, 4], from torchvision's own sampler."""
    from torchvision.models.detection._utils import BalancedPositiveNegativeSampler
    torch.manual_seed(seed)
    pos, neg = BalancedPositiveNegativeSampler(bs, frac)([row for row in labels])
    sampled = torch.zeros(labels.shape, dtype=torch.uint8)
    counts = torch.zeros(labels.shape[0], 4, dtype=torch.int32)
    for b, (p, n) in enumerate(zip(pos, neg)):
        sampled[b][p.bool()] = 1
        sampled[b][n.bool()] = 2
        counts[b] = torch.tensor([int((labels[b] >= 1).sum()), int((labels[b] == 0).sum()), int(p.sum()), int(n.sum())])
    return sampled, counts, pos, neg


def test_rpn_loss_static_matches_torchvision_cpu():
    from torchvision.models.detection.rpn import RegionProposalNetwork, RPNHead, AnchorGenerator
    g = torch.Generator().manual_seed(0)
    B, A = 3, 900
    rpn = RegionProposalNetwork(AnchorGenerator(), RPNHead(8, 3), 0.7, 0.3, 64, 0.5, dict(training=10, testing=10),
                                dict(training=10, testing=10), 0.7)
    labels = torch.zeros(B, A)
    u = torch.rand(B, A, generator=g)
    labels[u < 0.03] = 1
    labels[(u >= 0.03) & (u < 0.1)] = -1
    labels[2] = torch.where(torch.arange(A) < 40, labels[2], torch.full((A,), -1.0))      # fewer candidates than the batch size
    reg_t = torch.randn(B * A, 4, generator=g)
    reg_t[labels.reshape(-1) != 1] = float("-inf")          # what encode produces for an image without ground truth: never read
    obj = torch.randn(B * A, 1, generator=g, requires_grad=True)
    deltas = torch.randn(B * A, 4, generator=g, requires_grad=True)
    sampled, counts, pos, neg = _sampled_from_torchvision(labels, 64, 0.5, seed=3)
    torch.manual_seed(3)
    ref = rpn.compute_loss(obj, deltas, [row for row in labels], [t for t in reg_t.view(B, A, 4)])
    got = D.rpn_compute_loss_static(rpn, obj, deltas, labels, reg_t, sampled, counts)
    for a, b in zip(ref, got):
        assert torch.isfinite(b) and torch.allclose(a, b, rtol=1e-5, atol=1e-7), (float(a), float(b))
    ga = torch.autograd.grad(ref[0] + ref[1], (obj, deltas))
    gb = torch.autograd.grad(got[0] + got[1], (obj, deltas))
    for a, b in zip(ga, gb):
        assert torch.isfinite(b).all() and torch.allclose(a, b, rtol=1e-5, atol=1e-9)


def test_fastrcnn_loss_masked_matches_torchvision_cpu():
    from torchvision.models.detection.roi_heads import fastrcnn_loss
    g = torch.Generator().manual_seed(1)
    S, n, C = 96, 70, 3                                     # 96 rows, the last 26 are padding
    labels = torch.randint(0, C, (S,), generator=g)
    labels[n:] = -100
    reg_t = torch.randn(S, 4, generator=g)
    reg_t[labels <= 0] = float("nan")                       # background / padding targets are never read by torchvision
    logits = torch.randn(S, C, generator=g, requires_grad=True)
    box = torch.randn(S, 4 * C, generator=g, requires_grad=True)
    smp = D._StaticSamples(torch.zeros(S, 4), torch.zeros(S, dtype=torch.int64), labels, reg_t, torch.zeros(S, dtype=torch.int64),
                           torch.arange(S) < n, torch.tensor(n), torch.tensor([n]))
    ref = fastrcnn_loss(logits[:n], box[:n], [labels[:n]], [torch.nan_to_num(reg_t[:n])])
    got = D.fastrcnn_loss_masked(logits, box, smp)
    for a, b in zip(ref, got):
        assert torch.isfinite(b) and torch.allclose(a, b, rtol=1e-5, atol=1e-7), (float(a), float(b))
    ga = torch.autograd.grad(ref[0] + ref[1], (logits, box))
    gb = torch.autograd.grad(got[0] + got[1], (logits, box))
    for a, b in zip(ga, gb):
        assert torch.isfinite(b).all() and torch.allclose(a, b, rtol=1e-5, atol=1e-9) and bool((b[n:] == 0).all())


def test_sampled_rows_layout_cpu():
    sampled = torch.tensor([[0, 1, 0, 2, 2], [2, 0, 0, 0, 1], [0, 0, 0, 0, 0]], dtype=torch.uint8)
    counts = torch.tensor([[1, 4, 1, 2], [1, 4, 1, 1], [0, 0, 0, 0]], dtype=torch.int32)
    flat, valid, n = D._sampled_rows(sampled, counts, per_image=3)
    assert int(n) == 5 and flat[:5].tolist() == [1, 3, 4, 5, 9] and valid.tolist() == [True] * 5 + [False] * 4
