"""Forward time of the UNet and one DDIM step at the other BASELINE shapes (configs #1, #3-forward, #5) and for
target=joint (forward splat inside every step).  Prints one JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import FlowDiffuser  # noqa: E402
from opticalflowdiffusion_b200.config import compose  # noqa: E402
from opticalflowdiffusion_b200.datasets import synthetic_frames  # noqa: E402


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


res = {}
torch.set_grad_enabled(False)      # inference forward (with autograd on, unet(...) is the training forward)
torch.manual_seed(0)
flow_algo = FlowDiffuser(compose(["algorithm.target=flow", "algorithm.sampling_timesteps=50"]).algorithm).cuda()
joint_algo = FlowDiffuser(compose(["algorithm.target=joint", "algorithm.sampling_timesteps=50",
                                   "algorithm.zero_init=false"]).algorithm).cuda()
for name, B, H, W in (("64x128 b1", 1, 64, 128), ("368x768 b8", 8, 368, 768), ("436x1024 b8", 8, 436, 1024),
                      ("1024x2048 b2", 2, 1024, 2048)):
    cond = (2 * synthetic_frames(B, H, W, 0) - 1).cuda()
    x = torch.randn(B, 2, H, W, device="cuda")
    t = torch.full((B,), 500, device="cuda", dtype=torch.long)
    ms = timed(lambda: flow_algo.unet(x, cond, t))
    out = flow_algo.unet(x, cond, t)
    res[name] = {"unet_forward_ms": round(ms, 3), "ms_per_sample": round(ms / B, 3), "finite": bool(torch.isfinite(out).all())}
    if H <= 436:
        xj = torch.randn(B, 5, H, W, device="cuda")
        msj = timed(lambda: joint_algo._model(xj, cond, t))
        res[name]["joint_model_ms"] = round(msj, 3)
    torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
