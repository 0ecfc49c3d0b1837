"""GroupNorm + scale/shift + SiLU backward (fd_gn_silu_bwd = reduce + finalize + apply) at the training shapes:
CUDA-event time per call with the L2 flushed in between, and GB/s over the algorithmic bytes (reduce: h + da; apply:
h + da + dh; all bf16)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import _lib as L  # noqa: E402

lib = L.load(check_device=True)
out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (N, H, W, C) in [(8, 368, 768, 64), (8, 184, 384, 128), (8, 92, 192, 256), (8, 46, 96, 512)]:
    g = torch.Generator().manual_seed(C)
    h = (torch.randn(N, H, W, C, generator=g) * 2 + 0.5).to(torch.bfloat16).cuda()
    da = torch.randn(N, H, W, C, generator=g).to(torch.bfloat16).cuda()
    r = h.float().permute(0, 3, 1, 2).double().reshape(N, 8, -1)
    stats = torch.stack((r.sum(-1), (r * r).sum(-1)), -1).contiguous()
    gamma, beta = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    ss = (torch.randn(N, 2 * C, generator=g) * 0.5).cuda()
    dh = torch.empty_like(h)
    dgamma, dbeta, dbias = (torch.zeros(C, device="cuda") for _ in range(3))
    dss = torch.zeros(N, 2 * C, device="cuda")
    ws = torch.empty(lib.fd_gn_silu_bwd_workspace_floats(N, C), device="cuda")

    def call():
        L.check(lib.fd_gn_silu_bwd(L.ptr(h), L.ptr(da), L.ptr(stats), L.ptr(gamma), L.ptr(beta), L.ptr(ss), 2 * C, L.ptr(dh),
                                   L.ptr(dgamma), L.ptr(dbeta), L.ptr(dss), L.ptr(dbias), L.ptr(ws), N, H * W, C, 1e-5, L.stream()))

    for _ in range(3):
        call()
    ts = []
    for _ in range(int(os.environ.get("ITERS", 10))):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        call()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = 5 * h.numel() * 2
    out[f"{N}x{H}x{W}x{C}"] = {"us": round(ms * 1e3, 1), "GBps_5_passes": round(nbytes / ms / 1e6, 1)}
print(json.dumps(out, indent=1))
