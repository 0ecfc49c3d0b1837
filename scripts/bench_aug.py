"""Cost of the host-side Augmentor (augmentation.py:6-76 restated in flow_diffuser.Augmentor) and of training_step's
logging reductions at the training shape (batch 8, 368x768)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import FlowDiffuser  # noqa: E402
from opticalflowdiffusion_b200.config import compose  # noqa: E402
from opticalflowdiffusion_b200.datasets import synthetic_frames  # noqa: E402

B, H, W = 8, 368, 768
torch.manual_seed(0)
algo = FlowDiffuser(compose(["algorithm.target=flow"]).algorithm).cuda()
img, tgt = synthetic_frames(B, H, W, 1).cuda(), synthetic_frames(B, H, W, 2).cuda()
flow = torch.randn(B, 2, H, W, device="cuda") * 5


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print("preprocess aug=False: %.2f ms" % timed(lambda: algo.preprocess((img, tgt, flow), aug=False)))
print("preprocess aug=True : %.2f ms" % timed(lambda: algo.preprocess((img, tgt, flow), aug=True)))
first, cond, fl = algo.preprocess((img, tgt, flow), aug=False)
print("logging stats       : %.2f ms" % timed(lambda: {**algo._stats("train", "cond", cond), **algo._stats("train", "flow", fl)}))


def step():
    loss = algo.training_step((img, tgt, flow), 0)
    loss.backward()


print("training_step+backward (aug on, logging): %.2f ms" % timed(step, 5))
