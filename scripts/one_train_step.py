"""One training step (batch 8, 368x768 crops, target=flow; TARGET=joint / target selects the splat + pyramid-loss objectives)
after two warm-up steps: the command profiled by ncu for profiles/ (run with `--profile-from-start off`: only the last step is
inside cudaProfilerStart/Stop)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import FlowDiffuser  # noqa: E402
from opticalflowdiffusion_b200.config import compose  # noqa: E402
from opticalflowdiffusion_b200.datasets import synthetic_frames  # noqa: E402

B = int(os.environ.get("BATCH", 8))
H, W = int(os.environ.get("HEIGHT", 368)), int(os.environ.get("WIDTH", 768))
torch.manual_seed(0)
TARGET = os.environ.get("TARGET", "flow")
algo = FlowDiffuser(compose([f"algorithm.target={TARGET}", "algorithm.zero_init=false"]).algorithm).cuda()
opt = algo.configure_optimizers()
opt.max_grad_norm = 100.0
img, tgt = synthetic_frames(B, H, W, 1).cuda(), synthetic_frames(B, H, W, 2).cuda()
flow = torch.randn(B, 2, H, W, device="cuda") * 5
n = int(os.environ.get("STEPS", 3))
for i in range(n):
    if i == n - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    first, cond, fl = algo.preprocess((img, tgt, flow), aug=False)
    loss = algo.loss(first, cond, fl)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(loss.detach()))
