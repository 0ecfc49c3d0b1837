"""-m gpu: the training path -- UNet backward (UnetFunction), the loss gradient, the fused optimiser and
``FlowDiffuser.training_step`` -- against the reference-generated golden gradients (tests/golden/unet_flow_16x24.npz,
made by running the reference's ``p_losses(...).backward()``) and against autograd through the fp32 CPU oracle.

Tolerance (stated): activations and their gradients are bf16, accumulations fp32.  Per parameter tensor the
relative L2 error ||g - g_ref|| / ||g_ref|| must be <= 6e-2 and the cosine >= 0.998; the loss value within 2e-3."""
import numpy as np
import pytest
import torch

from oracle import flowdiff_oracle as O

pytestmark = pytest.mark.gpu

REL_L2 = 6e-2
COS = 0.998


def T(a):
    return torch.from_numpy(np.asarray(a))


def build_unet(seed, channels):
    from opticalflowdiffusion_b200.unet import Unet
    torch.manual_seed(int(seed))
    return Unet(64, channels=channels, out_dim=2)


def compare_grads(named_got, ref_sd, rel=REL_L2, cos=COS):
    worst = (0.0, None)
    fails = []
    for k, p in named_got:
        r = ref_sd[k].grad
        g = p.grad
        assert g is not None, k
        g = g.detach().float().cpu().flatten()
        r = r.float().flatten()
        nr = r.norm().item()
        if nr < 1e-10:
            assert g.norm().item() < 1e-6, k
            continue
        e = (g - r).norm().item() / nr
        c = torch.dot(g, r).item() / (g.norm().item() * nr + 1e-30)
        if e > worst[0]:
            worst = (e, k)
        if e > rel or c < cos:
            fails.append((k, round(e, 4), round(c, 5)))
    assert not fails, f"{len(fails)} parameter gradients out of tolerance (worst {worst}): {fails[:12]}"
    return worst


def oracle_grads(sd, fn):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = fn(sd)
    loss.backward()
    return sd, float(loss.detach())


def test_unet_backward_golden_16x24(golden):
    """loss = p_losses (target=flow) on the golden's inputs; gradients vs the reference's own backward and vs the oracle."""
    from opticalflowdiffusion_b200 import _lib
    g = golden("unet_flow_16x24")
    net = build_unet(g["seed"], 5)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    sched = O.make_schedule(1000)
    x0, cond, t, noise = T(g["x0"]), T(g["cond"]), T(g["t"]), T(g["noise"])
    ref_sd, ref_loss = oracle_grads(sd0, lambda sd: O.p_losses_flow(sd, sched, x0, cond, t, noise))
    x_t = O.q_sample(sched, x0, t, noise)
    out = net(x_t.cuda(), cond.cuda(), t.cuda())
    assert out.requires_grad and out.grad_fn is not None
    from opticalflowdiffusion_b200 import warp
    loss = warp.nan_mse(out, x0.cuda())
    assert abs(float(loss.detach()) - ref_loss) < 2e-3 * max(1.0, abs(ref_loss))
    n0 = _lib.load().fd_launch_count()
    loss.backward()
    assert _lib.load().fd_launch_count() - n0 > 300          # the backward ran on the library's kernels
    worst = compare_grads(net.named_parameters(), ref_sd)
    print("worst relative L2 gradient error", worst)
    # the reference's own numbers (golden) for three anchors
    gw = net.final_conv.weight.grad.cpu().numpy()
    np.testing.assert_allclose(gw, g["grad_final_conv_w"], rtol=0, atol=4e-2 * np.abs(g["grad_final_conv_w"]).max())
    gb = net.init_conv.bias.grad.cpu().numpy()
    np.testing.assert_allclose(gb, g["grad_init_conv_b"], rtol=0, atol=6e-2 * np.abs(g["grad_init_conv_b"]).max())
    s = float(net.mid_attn.fn.fn.to_qkv.weight.grad.double().abs().sum())
    assert abs(s - float(g["grad_mid_qkv_w_sum"])) < 5e-2 * float(g["grad_mid_qkv_w_sum"])


def test_unet_backward_vs_oracle_40x72():
    """40x72 (5x9 at the bottom level), B=2, every parameter gradient vs oracle autograd of sum(out * dout)."""
    net = build_unet(3, 5)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 40, 72
    x = torch.randn(B, 2, H, W, generator=g)
    cond = O.synthetic_frames(B, H, W, seed=6) * 2 - 1
    t = torch.tensor([812, 45])
    dout = torch.randn(B, 2, H, W, generator=g)
    ref_sd, _ = oracle_grads(sd0, lambda sd: (O.unet_forward(sd, x, cond, t) * dout).sum())
    out = net(x.cuda(), cond.cuda(), t.cuda())
    (out * dout.cuda()).sum().backward()
    worst = compare_grads(net.named_parameters(), ref_sd)
    print("worst relative L2 gradient error", worst)


def test_training_step_and_fused_adam():
    """FlowDiffuser.training_step -> backward -> FusedAdam: same loss as the oracle, p.grad on every parameter as
    views of one flat buffer, and the loss on a fixed batch decreases over a few steps."""
    from opticalflowdiffusion_b200 import FlowDiffuser
    from opticalflowdiffusion_b200.config import compose
    from opticalflowdiffusion_b200.optim import FusedAdam
    torch.manual_seed(0)
    cfg = compose(["algorithm.target=flow", "algorithm.lr=2e-4", "+algorithm.clipping=100"]).algorithm
    algo = FlowDiffuser(cfg).cuda()
    opt = algo.configure_optimizers()
    assert isinstance(opt, FusedAdam) and opt.max_grad_norm == 100.0
    B, H, W = 2, 32, 64
    img = O.synthetic_frames(B, H, W, seed=1).cuda()
    tgt = O.synthetic_frames(B, H, W, seed=2).cuda()
    flow = (torch.randn(B, 2, H, W, generator=torch.Generator().manual_seed(3)) * 5).cuda()
    first, cond, fl = algo.preprocess((img, tgt, flow), aug=False)
    t = torch.tensor([300, 700], device="cuda")
    noise = torch.randn(B, 2, H, W, generator=torch.Generator().manual_seed(4)).cuda()
    losses = []
    for step in range(6):
        loss = algo.model.p_losses(first, t, noise=noise, external_cond=cond)
        assert loss.requires_grad
        if step == 0:
            sd = {k[len("unet."):]: v.detach().cpu().clone() for k, v in algo.state_dict().items() if k.startswith("unet.")}
            ref = O.p_losses_flow(sd, O.make_schedule(1000), first.cpu(), cond.cpu(), t.cpu(), noise.cpu())
            assert abs(float(loss.detach()) - float(ref)) < 2e-3 * max(1.0, float(ref))
        loss.backward()
        grads = [p.grad for p in algo.unet.parameters()]
        assert all(gr is not None and torch.isfinite(gr).all() for gr in grads)
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses
    # the module API of the reference: training_step returns a scalar that carries the graph
    loss = algo.training_step((img, tgt, flow), 0)
    assert loss.dim() == 0 and loss.requires_grad
    loss.backward()
    assert algo.unet.final_conv.weight.grad is not None


def test_fused_adam_matches_torch_adam_on_unet():
    """One optimiser step on identical gradients: FusedAdam (flat, clip folded in) vs torch.optim.Adam + clip_grad_norm_."""
    from opticalflowdiffusion_b200.optim import FusedAdam
    net_a, net_b = build_unet(7, 5).cuda(), build_unet(7, 5).cuda()
    g = torch.Generator().manual_seed(8)
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        gr = torch.randn(pa.shape, generator=g).cuda() * 3.0
        pa.grad, pb.grad = gr.clone(), gr.clone()
    fa = FusedAdam(net_a.parameters(), lr=1e-3, weight_decay=1e-6, max_grad_norm=100.0)
    tb = torch.optim.Adam(net_b.parameters(), lr=1e-3, weight_decay=1e-6)
    torch.nn.utils.clip_grad_norm_(net_b.parameters(), 100.0)
    fa.step()
    tb.step()
    for (k, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6), k


@pytest.mark.parametrize("target", ["joint", "target"])
def test_training_step_through_the_splat(target):
    """The default config (target=joint, flow_diffuser.yaml:15): the UNet's flow drives the forward splat
    (UnetWithWarp, flow_diffuser.py:20-63) and the multi-scale loss (:893-983); gradients reach the UNet through
    fd_splat_flowgrad + UnetFunction.  Checks: finite loss / gradients on every parameter, and training reduces the
    loss on a fixed batch (zero_init off so that the predicted flow is not identically zero at step 0)."""
    from opticalflowdiffusion_b200 import FlowDiffuser
    from opticalflowdiffusion_b200.config import compose
    torch.manual_seed(0)
    cfg = compose([f"algorithm.target={target}", "algorithm.zero_init=false", "algorithm.lr=1e-4"]).algorithm
    algo = FlowDiffuser(cfg).cuda()
    opt = algo.configure_optimizers()
    opt.max_grad_norm = 100.0
    B, H, W = 2, 32, 64
    img = O.synthetic_frames(B, H, W, seed=1).cuda()
    tgt = O.synthetic_frames(B, H, W, seed=2).cuda()
    flow = (torch.randn(B, 2, H, W, generator=torch.Generator().manual_seed(3)) * 2).cuda()
    first, cond, fl = algo.preprocess((img, tgt, flow), aug=False)
    t = torch.tensor([100, 600], device="cuda")
    noise = torch.randn(first.shape, generator=torch.Generator().manual_seed(4)).cuda()
    losses = []
    for step in range(5):
        kw = dict(additional_tgt=fl, additional_weight=cfg.flow_weight) if target == "target" else {}
        loss = algo.model.p_losses(first, t, noise=noise, external_cond=cond, **kw)
        assert loss.requires_grad and torch.isfinite(loss)
        loss.backward()
        grads = [p.grad for p in algo.unet.parameters()]
        assert all(g is not None and torch.isfinite(g).all() for g in grads)
        if step == 0:
            assert sum(float(g.abs().sum()) for g in grads) > 0
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses


def test_experiment_train_task():
    """`experiment.tasks=[train]` through the runner (exp_base.py:178-214 without Lightning): loader -> training_step
    (with the Augmentor, flow_diffuser.py:219) -> backward -> clip + FusedAdam, two optimiser steps."""
    from opticalflowdiffusion_b200.config import compose
    from opticalflowdiffusion_b200.experiments import build_experiment
    cfg = compose(["algorithm.target=flow", "dataset.height=32", "dataset.width=32", "dataset.length=8",
                   "experiment.training.data.batch_size=2", "experiment.training.data.shuffle=false"])
    torch.manual_seed(0)
    exp = build_experiment(cfg, None, None)
    before = exp.algo.unet.init_conv.weight.detach().clone()
    out = exp.train(max_steps=2)
    assert out["steps"] == 2 and len(out["train/loss"]) == 2
    assert all(np.isfinite(v) for v in out["train/loss"])
    after = exp.algo.unet.init_conv.weight.detach().cpu()
    assert not torch.equal(before, after)                        # the fused Adam moved the parameters
    assert "train/loss" in exp.algo.logged


def test_backward_through_the_internal_padding():
    """Sizes that are not multiples of 8 are replicate-padded inside Unet.forward and the output is cropped: the
    backward of the ragged call must equal the backward of the explicitly padded call with zero gradient on the border."""
    import torch.nn.functional as F
    net = build_unet(11, 5).cuda()
    g = torch.Generator().manual_seed(12)
    B, H, W = 2, 36, 68                       # -> 40 x 72 internally, pad (2, 2, 2, 2)
    x = torch.randn(B, 2, H, W, generator=g).cuda()
    cond = torch.randn(B, 3, H, W, generator=g).cuda().clamp(-1, 1)
    t = torch.tensor([10, 900]).cuda()
    dout = torch.randn(B, 2, H, W, generator=g).cuda()
    out = net(x, cond, t)
    assert out.shape == (B, 2, H, W)
    (out * dout).sum().backward()
    ragged = [p.grad.clone() for p in net.parameters()]
    net.zero_grad(set_to_none=True)
    pad = (2, 2, 2, 2)
    xp, cp = F.pad(x, pad, mode="replicate"), F.pad(cond, pad, mode="replicate")
    outp = net(xp, cp, t)
    assert torch.equal(outp[:, :, 2:-2, 2:-2], out)
    (outp * F.pad(dout, pad)).sum().backward()
    padded = [p.grad.clone() for p in net.parameters()]
    net.zero_grad(set_to_none=True)
    (net(x, cond, t) * dout).sum().backward()                         # the ragged call again: run-to-run noise floor
    worst, noise = 0.0, 0.0
    for (k, p), r, q in zip(net.named_parameters(), ragged, padded):
        denom = q.norm().item() + 1e-12
        worst = max(worst, (q - r).norm().item() / denom)
        noise = max(noise, (p.grad - r).norm().item() / denom)
        # the activation-gradient chain is free of atomics (fixed-order reductions), so the two backward passes are the
        # same computation; only the fp32 atomics of the final parameter-gradient sums (wgrad, bias, gains) reorder
        assert (q - r).norm().item() <= 1e-4 * denom, (k, (q - r).norm().item() / denom)
        assert (p.grad - r).norm().item() <= 1e-4 * denom, (k, (p.grad - r).norm().item() / denom)
    print(f"padded-vs-ragged worst rel L2 {worst:.2e}; run-to-run {noise:.2e}")


def test_checkpoint_and_resume(tmp_path):
    """ModelCheckpoint(every_n_train_steps) + fit(ckpt_path=...) (exp_base.py:184-190,213): a run resumed from the step-2
    checkpoint reproduces the third step of the uninterrupted run (parameters and Adam moments restored)."""
    from opticalflowdiffusion_b200.config import compose
    from opticalflowdiffusion_b200.experiments import build_experiment
    ov = ["algorithm.target=flow", "algorithm.gpu_augment=true", "dataset.height=32", "dataset.width=32", "dataset.length=8",
          "experiment.training.data.batch_size=2", "experiment.training.data.shuffle=false",
          "experiment.training.checkpointing.every_n_train_steps=2", f"+output_dir={tmp_path}"]

    def run(ckpt, steps):
        import random
        random.seed(0)
        torch.manual_seed(0)
        exp = build_experiment(compose(ov), None, ckpt)
        exp.algo.preprocess = (lambda f: (lambda batch, aug=True: f(batch, aug=False)))(exp.algo.preprocess)   # no RNG in the data
        torch.manual_seed(123)                       # same t / noise draws in both runs from here on
        out = exp.train(max_steps=steps)
        return exp, out

    exp_a, out_a = run(None, 2)
    ck = tmp_path / "checkpoints" / "step_0000002.ckpt"
    assert ck.exists() and out_a["global_step"] == 2
    m_a = exp_a.algo.optimizers._m.clone()
    exp_b, out_b = run(str(ck), 0)
    assert out_b["global_step"] == 2
    for (k, a), b in zip(exp_a.algo.state_dict().items(), exp_b.algo.state_dict().values()):
        assert torch.equal(a.cpu(), b.cpu()), k
    # the restored moments land in the flat buffers at the first step
    exp_b.algo.optimizers._flatten()
    assert torch.equal(exp_b.algo.optimizers._m.cpu(), m_a.cpu()) and exp_b.algo.optimizers.step_count == 2
