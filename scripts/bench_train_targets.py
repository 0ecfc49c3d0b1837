"""Training step time per diffusion target (flow / joint / target) at batch 8, 368x768, aug off; with a phase split."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import FlowDiffuser  # noqa: E402
from opticalflowdiffusion_b200.config import compose  # noqa: E402
from opticalflowdiffusion_b200.datasets import synthetic_frames  # noqa: E402

B, H, W = int(os.environ.get("BATCH", 8)), 368, 768
img, tgt = synthetic_frames(B, H, W, 1).cuda(), synthetic_frames(B, H, W, 2).cuda()
flow = torch.randn(B, 2, H, W, device="cuda") * 5
for target in sys.argv[1:] or ["flow", "joint", "target"]:
    torch.manual_seed(0)
    algo = FlowDiffuser(compose([f"algorithm.target={target}", "algorithm.zero_init=false"]).algorithm).cuda()
    opt = algo.configure_optimizers()
    opt.max_grad_norm = 100.0

    def step():
        first, cond, fl = algo.preprocess((img, tgt, flow), aug=False)
        loss = algo.loss(first, cond, fl)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        loss = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    first, cond, fl = algo.preprocess((img, tgt, flow), aug=False)
    ev[1].record()
    loss = algo.loss(first, cond, fl)
    ev[2].record()
    loss.backward()
    ev[3].record()
    torch.cuda.synchronize()
    print(f"target={target}: {ms:.1f} ms/step ({B / ms * 1e3:.1f} samples/s)  preprocess {ev[0].elapsed_time(ev[1]):.1f}  "
          f"forward+loss {ev[1].elapsed_time(ev[2]):.1f}  backward {ev[2].elapsed_time(ev[3]):.1f} ms  loss {float(loss.detach()):.4f}")
    del algo, opt
    torch.cuda.empty_cache()
