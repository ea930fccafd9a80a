"""BASELINE.json configs 4 and 5 as timing side-notes (not the bench metric): RetinaNet train step, and inference-only
hallucination + detection at LLVIP resolution (1024x1280 -> S=300), batch limited by what one engine allocates."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallucidet_b200.train import HalluciDetTrainer
from oracle import step as ostep
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
res = {}

def timeit(fn, n=10, warm=4):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

# config 4: RetinaNet detector, train step, B=8, 512x640, S=640
tr = HalluciDetTrainer(detector_name="retinanet", size=640, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = ostep.synthetic_batch(8, 512, 640, seed=123, device=dev)
ms = timeit(lambda: float(tr.training_step(rgb, targets, ir, targets)["total"]))
res["config4_retinanet_train_B8"] = {"ms_per_step": ms, "images_per_s": 8 / ms * 1e3}
print("config 4 (RetinaNet train step, B=8):", res["config4_retinanet_train_B8"])
del tr; torch.cuda.empty_cache()

# config 5 family: inference at LLVIP resolution, S=300
B = int(os.environ.get("HD_INFER_BATCH", "8"))
tr = HalluciDetTrainer(detector_name="fasterrcnn", size=300, seed=123, device=dev, use_cuda_graph=True)
ir, rgb, targets = ostep.synthetic_batch(B, 1024, 1280, seed=5, device=dev)
ms = timeit(lambda: tr.test_step(rgb, targets, ir, targets), n=6, warm=3)
res[f"config5_inference_B{B}_1024x1280_S300"] = {"ms_per_batch": ms, "images_per_s": B / ms * 1e3,
                                                "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}
print("config 5 (inference, 1024x1280 -> S=300):", res[f"config5_inference_B{B}_1024x1280_S300"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/other_configs.json", "w"), indent=1)
