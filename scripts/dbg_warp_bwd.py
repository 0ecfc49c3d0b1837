import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import _lib
lib = _lib.load(check_device=True)
B, C, H, W = 2, 3, 24, 64
g = torch.Generator().manual_seed(5)
img = torch.rand(B, C, H, W, generator=g).cuda()
flow = (torch.randn(B, 2, H, W, generator=g) * 2).cuda()
gout = torch.randn(B, C, H, W, generator=g).cuda()
gi = torch.empty_like(img); gf = torch.empty_like(flow)
P, st = _lib.ptr, _lib.stream()
_lib.check(lib.fd_backwarp_bwd(P(img), P(flow), P(gout), P(gi), P(gf), B, C, H, W, st))
torch.cuda.synchronize()
print("ok", float(gi.abs().sum()), float(gf.abs().sum()))
