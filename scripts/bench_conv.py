"""Times fd_conv_igemm on the implicit-GEMM shapes of SURVEY.md appendix A (batch 8, 440x1024) and, with CUDNN=1, the library
bar for the same layers: torch.nn.functional.conv2d through cuDNN the way the reference runs it (fp32 NCHW tensors with TF32
allowed, main.py:79-80 / denoising_diffusion.py:114) and in the library's best configuration (bf16, channels_last), each with
cudnn.benchmark on.  The library legs include the concat of the two sources that the UNet's skip connections need
(torch.cat, denoising_diffusion.py:400-408) only when CAT=1; by default they get the already-concatenated tensor for free."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowdiffusion_b200 import _lib  # noqa: E402

SHAPES = [
    # name, C0, C1, Cout, H, W, k
    ("3x3 64->64 full", 64, 0, 64, 440, 1024, 3),
    ("3x3 128->64 full", 64, 64, 64, 440, 1024, 3),
    ("3x3 192->128 half", 128, 64, 128, 220, 512, 3),
    ("3x3 128->128 half", 128, 0, 128, 220, 512, 3),
    ("3x3 384->256 quarter", 256, 128, 256, 110, 256, 3),
    ("3x3 256->256 quarter", 256, 0, 256, 110, 256, 3),
    ("3x3 512->512 eighth", 512, 0, 512, 55, 128, 3),
    ("3x3 768->512 eighth", 512, 256, 512, 55, 128, 3),
    ("1x1 64->384 full", 64, 0, 384, 440, 1024, 1),
    ("1x1 128->64 full", 128, 0, 64, 440, 1024, 1),
]


def main():
    lib = _lib.load(check_device=True)
    N = int(os.environ.get("BATCH", 8))
    res = {}
    only = os.environ.get("ONLY")
    iters = int(os.environ.get("ITERS", 10))
    for name, c0, c1, cout, H, W, k in SHAPES:
        if only and only not in name:
            continue
        x0 = torch.randn(N, H, W, c0, device="cuda").to(torch.bfloat16)
        x1 = torch.randn(N, H, W, c1, device="cuda").to(torch.bfloat16) if c1 else None
        wp = (torch.randn(cout, k * k * (c0 + c1), device="cuda") * 0.02).to(torch.bfloat16)
        bias = torch.zeros(cout, device="cuda")
        out = torch.empty(N, H, W, cout, device="cuda", dtype=torch.bfloat16)
        gn = torch.zeros(N, 8, 2, device="cuda", dtype=torch.float64) if k == 3 else None
        P = _lib.ptr

        def fn():
            _lib.check(lib.fd_conv_igemm(P(x0), c0, P(x1), c1, P(wp), P(bias), None, P(out), P(gn), N, H, W, cout, k, k,
                                         k // 2, k // 2, 0, _lib.stream()))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) * 1e-3 / iters
        flops = 2.0 * N * H * W * cout * k * k * (c0 + c1)
        res[name] = {"ms": round(t * 1e3, 3), "TFLOPs": round(flops / t / 1e12, 1)}
        del x0, x1, out
        if os.environ.get("CUDNN"):
            import torch.nn.functional as F
            torch.backends.cudnn.benchmark = True
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
            for tag, dt, cl in (("cudnn_tf32_nchw", torch.float32, False), ("cudnn_bf16_nhwc", torch.bfloat16, True)):
                xs = [torch.randn(N, c, H, W, device="cuda", dtype=dt) for c in (c0, c1) if c]
                w = (torch.randn(cout, c0 + c1, k, k, device="cuda") * 0.02).to(dt)
                b = torch.zeros(cout, device="cuda", dtype=dt)
                if cl:
                    xs = [t_.contiguous(memory_format=torch.channels_last) for t_ in xs]
                    w = w.contiguous(memory_format=torch.channels_last)
                cat_inside = bool(os.environ.get("CAT")) and len(xs) > 1
                xin = xs[0] if len(xs) == 1 else torch.cat(xs, 1)

                def lib_fn():
                    return F.conv2d(torch.cat(xs, 1) if cat_inside else xin, w, b, padding=k // 2)
                for _ in range(3):
                    lib_fn()
                torch.cuda.synchronize()
                s.record()
                for _ in range(iters):
                    lib_fn()
                e.record()
                torch.cuda.synchronize()
                tl = s.elapsed_time(e) * 1e-3 / iters
                res[name][tag + "_ms"] = round(tl * 1e3, 3)
                res[name][tag + "_TFLOPs"] = round(flops / tl / 1e12, 1)
                del xs, xin, w
            res[name]["speedup_vs_tf32"] = round(res[name]["cudnn_tf32_nchw_ms"] / res[name]["ms"], 2)
            res[name]["speedup_vs_bf16"] = round(res[name]["cudnn_bf16_nhwc_ms"] / res[name]["ms"], 2)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
