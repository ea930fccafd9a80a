"""One eager train step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`), config 2 shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200.train import HalluciDetTrainer  # noqa: E402
from oracle import step as ostep  # noqa: E402

B = int(os.environ.get("HD_BATCH", "8"))
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True   # as bench.py
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=640, seed=123, device=dev, use_cuda_graph=False)
ir, rgb, targets = ostep.synthetic_batch(B, 512, 640, seed=123, device=dev)
for _ in range(3):
    tr.training_step(rgb, targets, ir, targets)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.training_step(rgb, targets, ir, targets)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
