"""oracle/randperm_cuda.py (numpy restatement of torch.randperm's CUDA algorithm + torchvision's balanced sampler) against golden
vectors produced by torch / torchvision themselves on a B200 (tests/golden/randperm_cuda.json, tests/golden/make_randperm_golden.py).
The product's device-side sampler is tested against torchvision directly on the GPU (tests/test_sampler_gpu.py); this pins the
algorithm it restates without one."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import randperm_cuda as R

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "randperm_cuda.json")))


@pytest.mark.parametrize("rec", GOLD["randperm"], ids=lambda r: f"seed{r['seed']}-off{r['offset']}-n{r['n']}")
def test_randperm_matches_torch_cuda(rec):
    perm, after = R.randperm(rec["n"], rec["seed"], rec["offset"])
    assert after == rec["offset_after"]
    assert perm[:512].tolist() == rec["head"]
    assert hashlib.sha256(perm.astype("int64").tobytes()).hexdigest() == rec["sha256"]
    assert sorted(perm.tolist()) == list(range(rec["n"]))


@pytest.mark.parametrize("i", range(len(GOLD["sampler"])))
def test_balanced_sampler_matches_torchvision_cuda(i):
    rec = GOLD["sampler"][i]
    sampled, after = R.balanced_sample(np.array(rec["labels"]), rec["bs"], rec["frac"], rec["seed"], rec["offset"])
    assert after == rec["offset_after"]
    assert sampled.tolist() == rec["sampled"]
