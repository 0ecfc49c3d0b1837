"""One UNet forward (batch 8, 436x1024) after a warm-up forward: the command profiled by ncu for profiles/."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import FlowDiffuser  # noqa: E402
from opticalflowdiffusion_b200.config import compose  # noqa: E402
from opticalflowdiffusion_b200.datasets import synthetic_frames  # noqa: E402

B = int(os.environ.get("BATCH", 8))
H, W = int(os.environ.get("HEIGHT", 436)), int(os.environ.get("WIDTH", 1024))
torch.set_grad_enabled(False)      # inference forward (with autograd on, unet(...) is the training forward)
torch.manual_seed(0)
algo = FlowDiffuser(compose(["algorithm.target=flow", "algorithm.sampling_timesteps=50"]).algorithm).cuda()
cond = (2 * synthetic_frames(B, H, W, 0) - 1).cuda()
x = torch.randn(B, 2, H, W, device="cuda")
t = torch.full((B,), 999, device="cuda", dtype=torch.long)
n_fwd = int(os.environ.get("FORWARDS", 2))
for i in range(n_fwd):
    if i == n_fwd - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()        # ncu --profile-from-start off: only the last (warm) forward is captured
    out = algo.unet(x, cond, t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(out.abs().mean()))
