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
