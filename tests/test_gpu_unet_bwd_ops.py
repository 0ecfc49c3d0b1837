"""-m gpu: the backward kernels of the UNet building blocks against torch autograd on fp32 restatements of
the reference ops (oracle/flowdiff_oracle.py), with bf16-rounded inputs.  Tolerances: activations' gradients are
stored in bf16 (2^-9 relative rounding) -> 1.5e-2 of the tensor's max; fp32 parameter gradients -> 5e-3."""
import pytest
import torch
import torch.nn.functional as F

from oracle import flowdiff_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def L():
    from opticalflowdiffusion_b200 import _lib
    _lib.load(check_device=True)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(BF)


def nchw(x):
    return x.permute(0, 3, 1, 2).float()


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-6)).item()


def bfr(x):
    return x.to(BF).float()


@pytest.mark.parametrize("C", [64, 128, 256, 512])
@pytest.mark.parametrize("use_ss", [True, False])
def test_gn_silu_bwd(L, C, use_ss):
    lib = L.load()
    g = torch.Generator().manual_seed(C + int(use_ss))
    N, H, W = 2, 9, 13
    x = bfr(torch.randn(N, C, H, W, generator=g) * 2 + 0.5).cuda().requires_grad_(True)
    gamma = torch.randn(C, generator=g).cuda().requires_grad_(True)
    beta = torch.randn(C, generator=g).cuda().requires_grad_(True)
    ss = (torch.randn(N, 3 * C, generator=g) * 0.5).cuda().requires_grad_(True)
    off = C // 2
    da = bfr(torch.randn(N, C, H, W, generator=g)).cuda()
    y = F.group_norm(x, 8, gamma, beta, eps=1e-5)
    if use_ss:
        sc, sh = ss[:, off:off + C], ss[:, off + C:off + 2 * C]
        y = y * (sc[:, :, None, None] + 1) + sh[:, :, None, None]
    F.silu(y).backward(da)
    r = x.detach().double().reshape(N, 8, -1)
    stats = torch.stack((r.sum(-1), (r * r).sum(-1)), -1).contiguous()
    xh, dah = nhwc(x.detach()), nhwc(da)
    dh = torch.empty_like(xh)
    dgamma, dbeta, dbias = (torch.zeros(C, device="cuda") for _ in range(3))
    dss = torch.zeros(N, 3 * C, device="cuda")
    ws = torch.empty(lib.fd_gn_silu_bwd_workspace_floats(N, C), device="cuda")
    L.check(lib.fd_gn_silu_bwd(L.ptr(xh), L.ptr(dah), L.ptr(stats), L.ptr(gamma), L.ptr(beta),
                               ss.data_ptr() + 4 * off if use_ss else None, 3 * C, L.ptr(dh), L.ptr(dgamma), L.ptr(dbeta),
                               dss.data_ptr() + 4 * off if use_ss else None, L.ptr(dbias), L.ptr(ws), N, H * W, C, 1e-5,
                               L.stream()))
    assert rel_err(nchw(dh), x.grad) < 1.5e-2
    assert rel_err(dgamma, gamma.grad) < 5e-3
    assert rel_err(dbeta, beta.grad) < 5e-3
    assert rel_err(dbias, x.grad.sum((0, 2, 3))) < 2e-2 or (dbias - x.grad.sum((0, 2, 3))).abs().max() < 2e-2
    if use_ss:
        assert rel_err(dss, ss.grad) < 5e-3


@pytest.mark.parametrize("C", [64, 128, 256, 512])
def test_chan_layernorm_bwd(L, C):
    lib = L.load()
    g = torch.Generator().manual_seed(C + 3)
    N, H, W = 2, 5, 11
    x = bfr(torch.randn(N, C, H, W, generator=g) * 3 + 1).cuda().requires_grad_(True)
    gain = torch.randn(1, C, 1, 1, generator=g).cuda().requires_grad_(True)
    dy = bfr(torch.randn(N, C, H, W, generator=g)).cuda()
    add = bfr(torch.randn(N, C, H, W, generator=g)).cuda()
    O._chan_layernorm(x, gain).backward(dy)
    xh, dyh, addh = nhwc(x.detach()), nhwc(dy), nhwc(add)
    for use_add in (False, True):
        dx = torch.empty_like(xh)
        dg = torch.zeros(C, device="cuda")
        L.check(lib.fd_chan_layernorm_bwd(L.ptr(xh), L.ptr(gain), L.ptr(dyh), L.ptr(addh) if use_add else None, L.ptr(dx),
                                          L.ptr(dg), N * H * W, C, 1e-5, L.stream()))
        assert rel_err(nchw(dx), x.grad + (add if use_add else 0)) < 1.5e-2
        assert rel_err(dg, gain.grad.flatten()) < 5e-3


def test_upsample_add_bias_final(L):
    lib = L.load()
    g = torch.Generator().manual_seed(5)
    dy = bfr(torch.randn(2, 64, 6, 10, generator=g)).cuda()
    dyh = nhwc(dy)
    dx = torch.empty(2, 3, 5, 64, device="cuda", dtype=BF)
    L.check(lib.fd_upsample2x_bwd(L.ptr(dyh), L.ptr(dx), 2, 3, 5, 64, L.stream()))
    ref = dy.reshape(2, 64, 3, 2, 5, 2).sum((3, 5))
    assert rel_err(nchw(dx), ref) < 1e-2
    # add
    a, b = nhwc(dy), nhwc(torch.randn(2, 64, 6, 10, generator=g).cuda())
    out = torch.empty_like(a)
    L.check(lib.fd_add_bf16(L.ptr(a), L.ptr(b), L.ptr(out), a.numel(), L.stream()))
    assert torch.equal(out, (a.float() + b.float()).to(BF))
    # bias gradient
    for C in (64, 384, 512):
        d = bfr(torch.randn(3, C, 7, 9, generator=g)).cuda()
        dh = nhwc(d)
        db = torch.zeros(C, device="cuda")
        L.check(lib.fd_bias_grad(L.ptr(dh), L.ptr(db), 3 * 63, C, L.stream()))
        assert rel_err(db, d.sum((0, 2, 3))) < 1e-4
    # final conv
    x = bfr(torch.randn(2, 64, 6, 10, generator=g)).cuda().requires_grad_(True)
    w = (torch.randn(2, 64, 1, 1, generator=g) * 0.1).cuda().requires_grad_(True)
    bias = torch.zeros(2, device="cuda", requires_grad=True)
    dout = torch.randn(2, 2, 6, 10, generator=g).cuda()
    F.conv2d(x, w, bias).backward(dout)
    xh = nhwc(x.detach())
    dxh = torch.empty_like(xh)
    dw, db = torch.zeros(2, 64, device="cuda"), torch.zeros(2, device="cuda")
    L.check(lib.fd_final_conv_bwd(L.ptr(xh), L.ptr(w), L.ptr(dout), L.ptr(dxh), L.ptr(dw), L.ptr(db), 2, 60, 64, 2, L.stream()))
    assert rel_err(nchw(dxh), x.grad) < 1e-2
    assert rel_err(dw, w.grad.reshape(2, 64)) < 1e-4
    assert rel_err(db, bias.grad) < 1e-4


@pytest.mark.parametrize("case", [(128, 64, 3, 3, 0, True), (64, 192, 3, 3, 0, False), (128, 256, 1, 1, 1, False),
                                  (64, 5, 7, 7, 2, False), (384, 64, 1, 1, 0, False)])
def test_weight_prep_bwd(L, case):
    """dgrad weight transposition and wgrad unpacking (+ weight-standardisation backward, :106-114)."""
    lib = L.load()
    Cout, Cin, KH, KW, kind, ws = case
    g = torch.Generator().manual_seed(Cout + Cin)
    w = torch.randn(Cout, Cin, KH, KW, generator=g).cuda().requires_grad_(True)
    Kp = 7 * 64 if kind == 2 else Cin * KH * KW
    packed = torch.empty(Cout, Kp, device="cuda", dtype=BF)
    L.check(lib.fd_prep_weight(L.ptr(w), L.ptr(packed), Cout, Cin, KH, KW, kind, int(ws), 1e-5, L.stream()))
    # an arbitrary upstream gradient in packed order
    gp = torch.randn(Cout, Kp, generator=g).cuda()
    # torch: packed = P(standardise(w)) with P a permutation -> pull gp back through the same permutation
    if ws:
        flat = w.reshape(Cout, -1)
        wt = ((w - flat.mean(1).reshape(-1, 1, 1, 1)) * (flat.var(1, unbiased=False).reshape(-1, 1, 1, 1) + 1e-5).rsqrt())
    else:
        wt = w
    if kind == 0:
        pk = wt.permute(0, 2, 3, 1).reshape(Cout, -1)
    elif kind == 1:
        C = Cin // 4
        pk = wt.reshape(Cout, C, 4).permute(0, 2, 1).reshape(Cout, 4 * C)
    else:
        pk = torch.zeros(Cout, 7, 64, device="cuda")
        pk[:, :, :KW * Cin] = wt.permute(0, 2, 3, 1).reshape(Cout, 7, KW * Cin)
        pk = pk.reshape(Cout, Kp)
    assert rel_err(packed.float(), pk.detach()) < 1e-2
    (pk * gp).sum().backward()
    dw = torch.zeros_like(w)
    L.check(lib.fd_prep_weight_bwd(L.ptr(gp), L.ptr(w), L.ptr(dw), Cout, Cin, KH, KW, kind, int(ws), 1e-5, L.stream()))
    assert rel_err(dw, w.grad) < 1e-4
    if kind != 2:
        T = KH * KW
        cin = Cin
        wd = torch.empty(cin, T * Cout, device="cuda", dtype=BF)
        if kind == 1:
            L.check(lib.fd_prep_weight_dgrad(L.ptr(packed), L.ptr(wd), Cout, cin, 1, L.stream()))
            assert torch.equal(wd, packed.t().contiguous())
        else:
            L.check(lib.fd_prep_weight_dgrad(L.ptr(packed), L.ptr(wd), Cout, cin, T, L.stream()))
            ref = packed.reshape(Cout, T, cin).flip(1).permute(2, 1, 0).reshape(cin, T * Cout)
            assert torch.equal(wd, ref.contiguous())


def test_time_path_bwd(L):
    """time_mlp (:319-324) and the ResnetBlock mlps (:193-196) backward from the two small dense-layer kernels."""
    lib = L.load()
    g = torch.Generator().manual_seed(4)
    B, J = 5, 384
    p = {"time_mlp.1.weight": torch.randn(256, 64, generator=g) * 0.1, "time_mlp.1.bias": torch.randn(256, generator=g),
         "time_mlp.3.weight": torch.randn(256, 256, generator=g) * 0.05, "time_mlp.3.bias": torch.randn(256, generator=g)}
    p = {k: v.requires_grad_(True) for k, v in p.items()}            # the oracle runs on the CPU
    wp_c = (torch.randn(J, 256, generator=g) * 0.05).requires_grad_(True)
    bp_c = torch.randn(J, generator=g).requires_grad_(True)
    t_c = torch.tensor([0, 1, 17, 500, 999])
    dss_c = torch.randn(B, J, generator=g)
    temb_ref = O.time_embedding(p, t_c)
    F.linear(F.silu(temb_ref), wp_c, bp_c).backward(dss_c)
    wp, bp, t, dss = wp_c.detach().cuda(), bp_c.detach().cuda(), t_c.cuda(), dss_c.cuda()
    temb, pe, pre = torch.empty(B, 256, device="cuda"), torch.empty(B, 64, device="cuda"), torch.empty(B, 256, device="cuda")
    w1, b1, w2, b2 = (p[k].detach().cuda() for k in ("time_mlp.1.weight", "time_mlp.1.bias", "time_mlp.3.weight", "time_mlp.3.bias"))
    st = L.stream()
    L.check(lib.fd_time_embed_save(L.ptr(t), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(temb), L.ptr(pe), L.ptr(pre),
                                   B, 64, 256, st))
    dwp, dbp = torch.zeros_like(wp), torch.zeros_like(bp)
    L.check(lib.fd_linear_bwd_w(L.ptr(dss), J, L.ptr(temb), 256, L.ptr(dwp), L.ptr(dbp), B, J, 256, 1, st))
    dtemb = torch.empty(B, 256, device="cuda")
    L.check(lib.fd_linear_bwd_x(L.ptr(dss), J, L.ptr(wp), L.ptr(temb), 256, L.ptr(dtemb), 256, B, J, 256, 1, st))
    dw2, db2 = torch.zeros_like(w2), torch.zeros_like(b2)
    L.check(lib.fd_linear_bwd_w(L.ptr(dtemb), 256, L.ptr(pre), 256, L.ptr(dw2), L.ptr(db2), B, 256, 256, 2, st))
    dpre = torch.empty(B, 256, device="cuda")
    L.check(lib.fd_linear_bwd_x(L.ptr(dtemb), 256, L.ptr(w2), L.ptr(pre), 256, L.ptr(dpre), 256, B, 256, 256, 2, st))
    dw1, db1 = torch.zeros_like(w1), torch.zeros_like(b1)
    L.check(lib.fd_linear_bwd_w(L.ptr(dpre), 256, L.ptr(pe), 64, L.ptr(dw1), L.ptr(db1), B, 256, 64, 0, st))
    for got, want in ((dwp, wp_c.grad), (dbp, bp_c.grad), (dw2, p["time_mlp.3.weight"].grad), (db2, p["time_mlp.3.bias"].grad),
                      (dw1, p["time_mlp.1.weight"].grad), (db1, p["time_mlp.1.bias"].grad)):
        assert rel_err(got.cpu(), want) < 2e-3, rel_err(got.cpu(), want)


def _split_heads(qkv):
    # (N, 384, H, W) -> q, k, v each (N, 4, 32, HW)
    n = qkv.shape[0]
    return [t.reshape(n, 4, 32, -1) for t in qkv.chunk(3, dim=1)]


@pytest.mark.parametrize("hw", [(6, 10), (17, 33), (40, 64)])
def test_linattn_bwd(L, hw):
    """LinearAttention core (:229-243) backward."""
    lib = L.load()
    H, W = hw
    N = 2
    g = torch.Generator().manual_seed(H)
    qkv = bfr(torch.randn(N, 384, H, W, generator=g)).cuda().requires_grad_(True)
    dout = bfr(torch.randn(N, 128, H, W, generator=g)).cuda()
    q, k, v = _split_heads(qkv)
    q = q.softmax(dim=-2) * 32 ** -0.5
    k = k.softmax(dim=-1)
    v = v / (H * W)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(N, 128, H, W)
    out.backward(dout)
    qh, dh = nhwc(qkv.detach()), nhwc(dout)
    dqkv = torch.empty_like(qh)
    ws = torch.empty(lib.fd_linattn_bwd_workspace_floats(N, H * W), device="cuda")
    L.check(lib.fd_linattn_bwd(L.ptr(qh), L.ptr(dh), L.ptr(dqkv), None, L.ptr(ws), N, H * W, L.stream()))
    got, want = nchw(dqkv), qkv.grad
    for name, sl in (("dq", slice(0, 128)), ("dk", slice(128, 256)), ("dv", slice(256, 384))):
        assert rel_err(got[:, sl], want[:, sl]) < 2e-2, (name, rel_err(got[:, sl], want[:, sl]))
    # with the statistics saved by the forward pass: same result, and the forward output matches
    att = torch.empty(N, H, W, 128, device="cuda", dtype=BF)
    stats = torch.empty(N, lib.fd_linattn_stats_floats(), device="cuda")
    wf = torch.empty(lib.fd_linattn_workspace_floats(N, H * W), device="cuda")
    L.check(lib.fd_linattn_save(L.ptr(qh), L.ptr(att), L.ptr(stats), L.ptr(wf), N, H * W, L.stream()))
    assert rel_err(nchw(att), out.detach()) < 1.5e-2
    dq2 = torch.empty_like(qh)
    L.check(lib.fd_linattn_bwd(L.ptr(qh), L.ptr(dh), L.ptr(dq2), L.ptr(stats), L.ptr(ws), N, H * W, L.stream()))
    assert rel_err(nchw(dq2), nchw(dqkv)) < 1e-2


@pytest.mark.parametrize("hw", [(4, 16), (9, 23), (24, 32)])
def test_attention_bwd(L, hw):
    """Attention core (:256-267) backward (flash-style, from the saved log-sum-exp)."""
    lib = L.load()
    H, W = hw
    N = 2
    g = torch.Generator().manual_seed(H + 1)
    qkv = bfr(torch.randn(N, 384, H, W, generator=g) * 1.5).cuda().requires_grad_(True)
    dout = bfr(torch.randn(N, 128, H, W, generator=g)).cuda()
    q, k, v = _split_heads(qkv)
    sim = torch.einsum("bhdi,bhdj->bhij", q * 32 ** -0.5, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhdj->bhid", attn, v)            # (N, 4, HW, 32)
    out = out.permute(0, 1, 3, 2).reshape(N, 128, H, W)
    out.backward(dout)
    qh, dh = nhwc(qkv.detach()), nhwc(dout)
    o = torch.empty(N, H, W, 128, device="cuda", dtype=BF)
    lse = torch.empty(N, 4, H * W, device="cuda")
    L.check(lib.fd_attention_lse(L.ptr(qh), L.ptr(o), L.ptr(lse), N, H * W, L.stream()))
    assert rel_err(nchw(o), out.detach()) < 1.5e-2
    ref_lse = torch.logsumexp(sim.detach(), dim=-1) * 1.4426950408889634
    assert (lse - ref_lse).abs().max() < 2e-2
    dqkv = torch.empty_like(qh)
    ws = torch.empty(lib.fd_attention_bwd_workspace_floats(N, H * W), device="cuda")
    L.check(lib.fd_attention_bwd(L.ptr(qh), L.ptr(o), L.ptr(dh), L.ptr(lse), L.ptr(dqkv), L.ptr(ws), N, H * W, L.stream()))
    got, want = nchw(dqkv), qkv.grad
    for name, sl in (("dq", slice(0, 128)), ("dk", slice(128, 256)), ("dv", slice(256, 384))):
        assert rel_err(got[:, sl], want[:, sl]) < 2.5e-2, (name, rel_err(got[:, sl], want[:, sl]))
    dq2 = torch.empty_like(qh)
    L.check(lib.fd_attention_bwd(L.ptr(qh), L.ptr(o), L.ptr(dh), L.ptr(lse), L.ptr(dq2), L.ptr(ws), N, H * W, L.stream()))
    assert torch.equal(dqkv, dq2)        # no atomics: bit-stable


def test_adam_clip(L):
    """fused clip + Adam against torch.optim.Adam (L2 weight decay) + clip_grad_norm_."""
    lib = L.load()
    g = torch.Generator().manual_seed(11)
    n = 100003
    p0 = torch.randn(n, generator=g).cuda()
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-2)
    p = torch.zeros(n + 1, device="cuda")[:n]
    p.copy_(p0)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        grad = (torch.randn(n, generator=g) * (5.0 if step == 2 else 0.01)).cuda()
        ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref], 100.0)
        opt.step()
        ss = torch.zeros(1, device="cuda")
        L.check(lib.fd_sumsq(L.ptr(grad), n, L.ptr(ss), L.stream()))
        assert abs(ss.item() - grad.double().pow(2).sum().item()) < 1e-4 * ss.item()
        L.check(lib.fd_adam_step(L.ptr(p), L.ptr(grad), L.ptr(m), L.ptr(v), n, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, L.ptr(ss),
                                 100.0, 1.0, L.stream()))
        assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-6), (p - ref.detach()).abs().max()


def test_weight_prep_batched_equals_per_layer(L):
    """fd_prep_weight_batch / _dgrad_batch / _bwd_batch (one launch over a device table of layers) against the per-layer
    entry points: identical arithmetic per block, so the results must be bit-equal."""
    lib = L.load()
    g = torch.Generator().manual_seed(11)
    layers = [  # Cout, Cin, KH, KW, kind, standardize
        (64, 64, 3, 3, 0, 1), (128, 192, 3, 3, 0, 1), (64, 9, 7, 7, 2, 0), (128, 256, 1, 1, 1, 0), (384, 64, 1, 1, 0, 0),
        (40, 72, 3, 3, 0, 1),
    ]
    ws = [torch.randn(co, ci, kh, kw, generator=g).cuda() for co, ci, kh, kw, _, _ in layers]

    def kp(l):
        co, ci, kh, kw, kind, _ = l
        return kh * 64 if kind == 2 else ci * kh * kw

    # --- forward packing
    ref = [torch.empty(l[0], kp(l), device="cuda", dtype=torch.bfloat16) for l in layers]
    out = [torch.zeros_like(r) for r in ref]
    for l, w, r in zip(layers, ws, ref):
        L.check(lib.fd_prep_weight(L.ptr(w), L.ptr(r), l[0], l[1], l[2], l[3], l[4], l[5], 1e-5, L.stream()))
    recs = [[w.data_ptr(), o.data_ptr(), 0, l[0], l[1], l[2], l[3], l[4] | (l[5] << 8)] for l, w, o in zip(layers, ws, out)]
    blk = [0]
    for l in layers:
        blk.append(blk[-1] + l[0])
    table = torch.tensor(recs, dtype=torch.int64).cuda()
    blk_t = torch.tensor(blk, dtype=torch.int32).cuda()
    L.check(lib.fd_prep_weight_batch(L.ptr(table), L.ptr(blk_t), len(layers), blk[-1], 1e-5, L.stream()))
    torch.cuda.synchronize()
    for r, o in zip(ref, out):
        assert torch.equal(r, o)

    # --- dgrad weights (kinds 0 / 1 only)
    dl = [(l, r) for l, r in zip(layers, ref) if l[4] != 2]
    dref, dout, drecs, dblk = [], [], [], [0]
    for l, packed in dl:
        co, ci, kh, kw, kind, _ = l
        taps = 1 if kind == 1 else kh * kw
        cin = packed.shape[1] // taps
        a = torch.empty(cin, taps * co, device="cuda", dtype=torch.bfloat16)
        b = torch.zeros_like(a)
        L.check(lib.fd_prep_weight_dgrad(L.ptr(packed), L.ptr(a), co, cin, taps, L.stream()))
        dref.append(a)
        dout.append(b)
        drecs.append([packed.data_ptr(), b.data_ptr(), 0, co, cin, taps, 0, 0])
        dblk.append(dblk[-1] + ((cin + 31) // 32) * ((co + 31) // 32) * taps)
    dtable, dblk_t = torch.tensor(drecs, dtype=torch.int64).cuda(), torch.tensor(dblk, dtype=torch.int32).cuda()
    L.check(lib.fd_prep_weight_dgrad_batch(L.ptr(dtable), L.ptr(dblk_t), len(drecs), dblk[-1], L.stream()))
    torch.cuda.synchronize()
    for a, b in zip(dref, dout):
        assert torch.equal(a, b)

    # --- wgrad unpacking + weight-standardisation backward
    gs = [torch.randn(l[0], kp(l), generator=g).cuda() for l in layers]
    bref = [torch.full_like(w, 0.25) for w in ws]
    bout = [torch.full_like(w, 0.25) for w in ws]
    for l, gp, w, d in zip(layers, gs, ws, bref):
        L.check(lib.fd_prep_weight_bwd(L.ptr(gp), L.ptr(w), L.ptr(d), l[0], l[1], l[2], l[3], l[4], l[5], 1e-5, L.stream()))
    brecs = [[gp.data_ptr(), w.data_ptr(), d.data_ptr(), l[0], l[1], l[2], l[3], l[4] | (l[5] << 8)]
             for l, gp, w, d in zip(layers, gs, ws, bout)]
    btable = torch.tensor(brecs, dtype=torch.int64).cuda()
    L.check(lib.fd_prep_weight_bwd_batch(L.ptr(btable), L.ptr(blk_t), len(layers), blk[-1], 1e-5, L.stream()))
    torch.cuda.synchronize()
    for a, b in zip(bref, bout):
        assert torch.equal(a, b)
