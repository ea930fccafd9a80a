"""Golden vectors for oracle/randperm_cuda.py: outputs of torch.randperm(n, device="cuda") and of torchvision's
BalancedPositiveNegativeSampler on the CUDA generator, for fixed (seed, offset) states.  Needs a CUDA device:
    python tests/golden/make_randperm_golden.py        (writes tests/golden/randperm_cuda.json)
Large permutations are stored as their first 512 entries + the SHA-256 of the whole int64 array."""
import hashlib
import json
import os

import torch

out = {"torch": torch.__version__, "device": torch.cuda.get_device_name(0), "randperm": [], "sampler": []}
gen = None
torch.cuda.init()
gen = torch.cuda.default_generators[0]
STATES = [(7, 0), (123456789012345, 4096), ((1 << 40) + 5, 8)]
for seed, offset in STATES:
    for n in (1, 2, 3, 5, 9, 17, 33, 64, 300, 2000, 4096, 30083, 30084, 100000):
        gen.manual_seed(seed)
        gen.set_offset(offset)
        p = torch.randperm(n, device="cuda")
        after = int(gen.get_offset())
        a = p.cpu().numpy().astype("int64")
        rec = {"seed": seed, "offset": offset, "n": n, "offset_after": after, "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
               "head": a[:512].tolist()}
        out["randperm"].append(rec)
from torchvision.models.detection._utils import BalancedPositiveNegativeSampler
g = torch.Generator().manual_seed(11)
for (B, N, p_pos, p_ign, bs, frac) in ((3, 500, 0.05, 0.1, 64, 0.5), (2, 40, 0.4, 0.2, 16, 0.25), (2, 3000, 0.01, 0.0, 256, 0.5)):
    u = torch.rand(B, N, generator=g)
    lab = torch.zeros(B, N)
    lab[u < p_pos] = 1
    lab[(u >= p_pos) & (u < p_pos + p_ign)] = -1
    seed, offset = 99, 16
    gen.manual_seed(seed)
    gen.set_offset(offset)
    pos, neg = BalancedPositiveNegativeSampler(bs, frac)([row.cuda() for row in lab])
    sampled = torch.zeros(B, N, dtype=torch.uint8)
    for b in range(B):
        sampled[b][pos[b].bool().cpu()] = 1
        sampled[b][neg[b].bool().cpu()] = 2
    out["sampler"].append({"seed": seed, "offset": offset, "labels": lab.to(torch.int8).tolist(), "bs": bs, "frac": frac,
                           "sampled": sampled.tolist(), "offset_after": int(gen.get_offset())})
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "randperm_cuda.json")
json.dump(out, open(path, "w"))
print("wrote", path, os.path.getsize(path), "bytes")
