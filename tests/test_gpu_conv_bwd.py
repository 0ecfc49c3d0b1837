"""-m gpu: data / weight gradients of the convolution engine against torch autograd (fp32 conv2d on the same
bf16-rounded operands).  dgrad = fd_conv_igemm(_ex) on flipped / transposed packed weights; wgrad = fd_conv_wgrad."""
import pytest
import torch
import torch.nn.functional as F

from test_gpu_conv import nhwc_bf16, pack_w, close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from opticalflowdiffusion_b200 import _lib
    _lib.load(check_device=True)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib


def unshuffle(x):
    b, c, H, W = x.shape
    return x.reshape(b, c, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(b, c * 4, H // 2, W // 2)


def torch_grads(xs, w, dy, pad, mode):
    x = torch.cat([t.to(torch.bfloat16).float() for t in xs], 1).requires_grad_(True)
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    y = F.conv2d(unshuffle(x) if mode == 1 else x, wq, None, padding=pad)
    y.backward(dy.to(torch.bfloat16).float())
    return x.grad, wq.grad


WG_CASES = [
    # N, Cins, Cout, H, W, k, pad, mode
    (2, (64,), 64, 16, 128, (3, 3), (1, 1), 0),
    (1, (64,), 64, 5, 70, (3, 3), (1, 1), 0),
    (2, (64, 64), 128, 12, 24, (3, 3), (1, 1), 0),
    (2, (128, 64), 128, 9, 16, (3, 3), (1, 1), 0),
    (2, (64,), 384, 10, 12, (1, 1), (0, 0), 0),
    (2, (128,), 64, 10, 13, (1, 1), (0, 0), 0),
    (1, (64,), 64, 20, 40, (7, 1), (3, 0), 0),
    (2, (512, 256), 512, 4, 6, (3, 3), (1, 1), 0),
    (1, (256,), 256, 33, 70, (3, 3), (1, 1), 0),
    (2, (64,), 64, 110, 256, (3, 3), (1, 1), 0),           # rolling-strip wgrad: two column blocks
    (1, (64,), 64, 37, 130, (3, 3), (1, 1), 0),            # ragged W: second column block is 2 pixels wide
    (2, (64, 64), 64, 21, 200, (3, 3), (1, 1), 0),         # two sources -> two ci chunks
    (1, (128,), 128, 19, 96, (3, 3), (1, 1), 0),           # 2 x 2 block pairs, W < 128
    (3, (192,), 128, 7, 64, (3, 3), (1, 1), 0),            # more block pairs than rows per CTA range
    (2, (64,), 128, 8, 12, (1, 1), (0, 0), 1),
    (1, (128,), 256, 16, 32, (1, 1), (0, 0), 1),
]


def make(case):
    N, cins, Cout, H, W, k, pad, mode = case
    g = torch.Generator().manual_seed(sum(cins) + H * 7 + W)
    s = 2 if mode == 1 else 1
    xs = [torch.randn(N, c, H * s, W * s, generator=g).cuda() for c in cins]
    cin = sum(cins) * (4 if mode == 1 else 1)
    w = (torch.randn(Cout, cin, k[0], k[1], generator=g) / (cin * k[0] * k[1]) ** 0.5).cuda()
    dy = torch.randn(N, Cout, H, W, generator=g).cuda()
    return xs, w, dy


@pytest.mark.parametrize("case", WG_CASES)
def test_conv_wgrad(L, case):
    N, cins, Cout, H, W, k, pad, mode = case
    xs, w, dy = make(case)
    lib = L.load()
    srcs = [nhwc_bf16(x) for x in xs]
    dyp = nhwc_bf16(dy)
    K = w[0].numel()
    dw = torch.zeros(Cout, K, device="cuda", dtype=torch.float32)
    L.check(lib.fd_conv_wgrad(L.ptr(srcs[0]), cins[0], L.ptr(srcs[1]) if len(srcs) > 1 else None,
                              cins[1] if len(cins) > 1 else 0, L.ptr(dyp), L.ptr(dw), N, H, W, Cout, k[0], k[1], pad[0],
                              pad[1], mode, L.stream()))
    torch.cuda.synchronize()
    _, gw = torch_grads(xs, w, dy, pad, mode)
    if mode == 0:
        ref = gw.permute(0, 2, 3, 1).reshape(Cout, -1)
    else:
        C = cins[0]
        ref = gw.reshape(Cout, C, 4).permute(0, 2, 1).reshape(Cout, 4 * C)
    err = (dw - ref).abs()
    scale = ref.abs().max().item()
    assert err.max().item() <= 2e-3 * scale + 1e-3, f"max err {err.max().item()} scale {scale}"


DG_CASES = [c for c in WG_CASES if c[5] != (7, 1)]


@pytest.mark.parametrize("case", DG_CASES)
def test_conv_dgrad(L, case):
    N, cins, Cout, H, W, k, pad, mode = case
    xs, w, dy = make(case)
    lib = L.load()
    dyp = nhwc_bf16(dy)
    gx, _ = torch_grads(xs, w, dy, pad, mode)
    wf = pack_w(w) if mode == 0 else None
    if mode == 0:
        taps, cin = k[0] * k[1], sum(cins)
        # wd[ci][(T-1-tap)*Cout + co] = wf[co][tap*Cin + ci]
        wd = wf.reshape(Cout, taps, cin).flip(1).permute(2, 1, 0).reshape(cin, taps * Cout).contiguous()
        off = 0
        for c in cins:
            out = torch.empty(N, H, W, c, device="cuda", dtype=torch.bfloat16)
            L.check(lib.fd_conv_igemm_ex(L.ptr(dyp), Cout, None, 0, L.ptr(wd[off:off + c]), None, None, L.ptr(out), None,
                                         N, H, W, c, k[0], k[1], pad[0], pad[1], 0, 0, L.stream()))
            torch.cuda.synchronize()
            close(out.permute(0, 3, 1, 2).float(), gx[:, off:off + c])
            off += c
    else:
        C = cins[0]
        wp = w.reshape(Cout, C, 4).permute(0, 2, 1).reshape(Cout, 4 * C).to(torch.bfloat16)
        wd = wp.t().contiguous()                     # [(p*C + c)][co]
        out = torch.empty(N, 2 * H, 2 * W, C, device="cuda", dtype=torch.bfloat16)
        L.check(lib.fd_conv_igemm_ex(L.ptr(dyp), Cout, None, 0, L.ptr(wd), None, None, L.ptr(out), None, N, H, W, 4 * C,
                                     1, 1, 0, 0, 0, 1, L.stream()))
        torch.cuda.synchronize()
        close(out.permute(0, 3, 1, 2).float(), gx)
