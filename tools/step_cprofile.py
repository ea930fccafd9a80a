"""cProfile of the host side of the config-2 train step (where the Python / launch time goes)."""
import cProfile, os, pstats, sys, io
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200.train import HalluciDetTrainer
from oracle import step as ostep
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = ostep.synthetic_batch(8, 512, 640, seed=123, device=dev)
for _ in range(5):
    tr.training_step(rgb, targets, ir, targets)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    out = tr.training_step(rgb, targets, ir, targets)
    float(out["total"])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:6000])
