"""FlowLearner training step (algorithm=flow_learner, 8f row N1) at the reference's default image_size 128, batch 16:
step time with a phase split, and the CPU oracle on the host cores for the same objective (one sample)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import FlowLearner, _lib  # noqa: E402
from opticalflowdiffusion_b200.config import compose  # noqa: E402
from opticalflowdiffusion_b200.datasets import synthetic_frames  # noqa: E402

B, H, W = int(os.environ.get("BATCH", 16)), int(os.environ.get("SIZE", 128)), int(os.environ.get("SIZE", 128))
torch.manual_seed(0)
algo = FlowLearner(compose(["algorithm=flow_learner", "algorithm.zero_init=false"]).algorithm).cuda()
opt = algo.configure_optimizers()
opt.max_grad_norm = 100.0
img, tgt = synthetic_frames(B, H, W, 1).cuda(), synthetic_frames(B, H, W, 2).cuda()
flow = torch.randn(B, 2, H, W, device="cuda") * 3
lib = _lib.load()


def step():
    t, c, f = algo.preprocess((img, tgt, flow), aug=True)
    loss = algo.loss(t, c, f)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
n0 = lib.fd_launch_count()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    loss = step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / n * 1e3
launches = (lib.fd_launch_count() - n0) / n
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
t, c, f = algo.preprocess((img, tgt, flow), aug=False)
ev[0].record()
fp, ww = algo._predict(c)
ev[1].record()
loss = algo.objective(c[:, :3].contiguous(), t, fp, ww)
ev[2].record()
loss.backward()
ev[3].record()
torch.cuda.synchronize()
print(f"FlowLearner b{B} {H}x{W}: {ms:.1f} ms/step ({B / ms * 1e3:.1f} samples/s), {launches:.0f} library launches/step; "
      f"UNet forward {ev[0].elapsed_time(ev[1]):.1f} ms, 832-term objective {ev[1].elapsed_time(ev[2]):.1f} ms, "
      f"backward {ev[2].elapsed_time(ev[3]):.1f} ms, loss {float(loss.detach()):.4f}")
if os.environ.get("CPU", "1") != "0":
    from oracle import flowdiff_oracle as O
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in algo.unet.model.state_dict().items()}
    tc, cc = t[:1].cpu(), c[:1].cpu()
    t0 = time.perf_counter()
    l = O.flow_learner_loss(sd, tc, cc)
    l.backward()
    sec = time.perf_counter() - t0
    print(f"CPU oracle (port), {torch.get_num_threads()} threads: {sec:.1f} s for ONE sample forward + backward "
          f"-> {1.0 / sec:.3f} samples/s")
