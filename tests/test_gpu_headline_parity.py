"""-m gpu: parity AT THE BASELINE SHAPES ("parity exists only at toy shapes").

Every shape-dispatched path of the conv engine (148-range strip walks, CTA pairs with two M-tiles, fused
GroupNorm inputs with padding rows, the deep TMA rings) only becomes the majority at these sizes, and the
headline number is bf16 against an fp32 reference, so the tolerance evidence has to be taken here:

  (a) teacher-forced x0 prediction at 436x1024 (UNet on 440x1024) at EVERY one of the 50 DDIM steps:
      the oracle's own x_t is the input, so bf16 error is not compounded;
  (b) free-running DDIM-50 at 436x1024: end-point error in pixels (x flow_max = 20) between the CUDA
      trajectory and the fp32 oracle trajectory, and |EPE_new - EPE_ref| against a synthetic ground truth;
  (c) one 368x768 training step (BASELINE configs[2] crop): loss value and all 276 parameter gradients
      against autograd through the oracle;
  (d) BASELINE configs[3] at the full 8x2x436x1024: backward warp bit-exact, photometric / EPE forward
      and both gradients against the oracle.

Tolerances are written next to each assert.  The measured numbers are also dumped to
``gpurun_out/parity_headline.json`` so DESIGN.md can quote them.
"""
import json
import os
import time

import numpy as np
import pytest
import torch

from oracle import flowdiff_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOW_MAX = 20.0


def _record(key, value):
    path = os.path.join(ROOT, "gpurun_out", "parity_headline.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[key] = value
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass
    print(key, json.dumps(value))


def _all_threads():
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        pass


def _epe(a, b):
    return torch.sqrt(((a - b) * FLOW_MAX).pow(2).sum(1)).mean().item()


def _build(seed, sampling_timesteps=50):
    from opticalflowdiffusion_b200 import FlowDiffuser
    from opticalflowdiffusion_b200.config import compose
    torch.manual_seed(seed)
    algo = FlowDiffuser(compose(["algorithm.target=flow", f"algorithm.sampling_timesteps={sampling_timesteps}",
                                 "algorithm.return_all_timesteps=false"]).algorithm)
    sd = {k: v.detach().clone() for k, v in algo.unet.state_dict().items()}
    return algo.cuda(), sd


def test_ddim50_436x1024_teacher_forced_and_free_running():
    """(a) + (b).  One oracle DDIM-50 trajectory at 436x1024, batch 1 (50 fp32 CPU forwards), reused for both checks.

    Tolerances: teacher-forced RAW prediction (before the sampler's clamp to [-1, 1]; it reaches +-2.8 at the late steps
    with random-init weights) max |err| <= 4 % of max(1, max |ref|) -- the per-tensor rule of tests/test_gpu_unet.py -- and
    mean |err| <= 5e-3 at every step; the CLAMPED prediction the sampler actually uses max |err| <= 6e-2 (3 % of its range;
    measured 5.0e-2 at a single pixel of step t = 19, 2.3e-2 at t = 499, 1.6e-2 at t = 999); free-running final
    flow EPE(new, ref) <= 0.25 px of a +-20 px range and |EPE_new - EPE_ref| <= 0.1 px against the synthetic ground truth."""
    _all_threads()
    H, W, S = 436, 1024, 50
    algo, sd = _build(0, S)
    sched = O.make_schedule(1000)
    cond = O.synthetic_frames(1, H, W, seed=100) * 2 - 1
    x_T = torch.randn(1, 2, H, W, generator=torch.Generator().manual_seed(1234))
    cond_p, pad = O.replicate_pad_to_multiple(cond)
    xT_p, _ = O.replicate_pad_to_multiple(x_T)
    top, left = pad[2], pad[0]

    def crop(v):
        return v[..., top:top + H, left:left + W]

    # the oracle has no internal padding: it runs on the replicate-padded 440x1024 frames and is cropped back, which is
    # what the CUDA path does inside Unet.forward (future/raft_utils.py:7-25 semantics); the padded border of x_t is
    # re-padded from the cropped state every step so both see the same input
    times = O.ddim_times(1000, S)
    x = x_T
    ref_states, ref_x0 = [x_T], []
    t0 = time.perf_counter()
    with torch.no_grad():
        for tm, tn in zip(times[:-1], times[1:]):
            xp, _ = O.replicate_pad_to_multiple(x)
            out = crop(O.unet_forward(sd, xp, cond_p, torch.full((1,), tm, dtype=torch.long)))
            x, x0 = O.ddim_update(sched, x, out, tm, tn)
            ref_states.append(x)
            ref_x0.append(out)
    cpu_s = time.perf_counter() - t0
    del xT_p

    # (a) teacher-forced, every step
    cond_d = cond.cuda()
    worst_rel, worst_mean, worst_clamped, rows = 0.0, 0.0, 0.0, []
    with torch.no_grad():
        for i, tm in enumerate(times[:-1]):
            out = algo.unet(ref_states[i].cuda(), cond_d, torch.full((1,), tm, device="cuda", dtype=torch.long)).cpu()
            err = (out - ref_x0[i]).abs()
            cerr = (out.clamp(-1, 1) - ref_x0[i].clamp(-1, 1)).abs().max().item()
            scale = max(1.0, ref_x0[i].abs().max().item())
            rows.append((tm, err.max().item(), err.mean().item(), ref_x0[i].abs().max().item(), cerr))
            worst_rel, worst_mean = max(worst_rel, rows[-1][1] / scale), max(worst_mean, rows[-1][2])
            worst_clamped = max(worst_clamped, cerr)
    _record("teacher_forced_436x1024", {"steps": S, "worst_max_err_over_scale": worst_rel, "worst_mean_abs_err": worst_mean,
                                        "worst_clamped_max_abs_err": worst_clamped,
                                        "t999": rows[0][1:], "t499": rows[25][1:], "t19": rows[-1][1:],
                                        "cols": "max|err|, mean|err|, max|ref|, max|err| after the sampler's clamp",
                                        "oracle_cpu_seconds": cpu_s})

    # (b) free-running DDIM-50 through the public sampler (eager and CUDA-graph replay)
    out = algo.model.sample(1, external_cond=cond_d, x_T=x_T)
    ref = ref_states[-1]
    epe = _epe(out.cpu(), ref)
    g = torch.Generator().manual_seed(5)
    gt = torch.clamp(torch.randn(1, 2, H, W, generator=g) * 0.2, -1, 1)        # synthetic ground truth, sigma = 4 px
    epe_new, epe_ref = _epe(out.cpu(), gt), _epe(ref, gt)
    graphed = algo.model.sample(1, external_cond=cond_d, x_T=x_T, use_cuda_graph=True)
    _record("ddim50_436x1024", {"epe_px_new_vs_ref": epe, "epe_px_new_vs_gt": epe_new, "epe_px_ref_vs_gt": epe_ref,
                                "max_abs_diff_px": (out.cpu() - ref).abs().max().item() * FLOW_MAX,
                                "graph_vs_eager_max_abs": (graphed - out).abs().max().item()})
    assert worst_rel <= 4e-2, rows
    assert worst_mean <= 5e-3, rows
    assert worst_clamped <= 6e-2, rows
    assert epe <= 0.25, f"EPE(CUDA DDIM-50, oracle DDIM-50) = {epe} px"
    assert abs(epe_new - epe_ref) <= 0.1, (epe_new, epe_ref)
    assert (graphed - out).abs().max().item() < 5e-3


def test_training_step_368x768_gradients_vs_oracle():
    """(c).  p_losses (target=flow) at one 368x768 crop: loss within 2e-3 relative; per parameter tensor
    ||g - g_ref|| / ||g_ref|| <= 6e-2 and cosine >= 0.998 (the tolerances of tests/test_gpu_train.py)."""
    _all_threads()
    H, W = 368, 768
    algo, sd = _build(1)
    sched = O.make_schedule(1000)
    g = torch.Generator().manual_seed(11)
    cond = O.synthetic_frames(1, H, W, seed=200) * 2 - 1
    x0 = torch.clamp(torch.randn(1, 2, H, W, generator=g) * 0.25, -1, 1)
    noise = torch.randn(1, 2, H, W, generator=g)
    t = torch.tensor([417])
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_loss = O.p_losses_flow(ref_sd, sched, x0, cond, t, noise)
    ref_loss.backward()
    loss = algo.model.p_losses(x0.cuda(), t.cuda(), noise=noise.cuda(), external_cond=cond.cuda())
    loss.backward()
    rel = abs(float(loss.detach()) - float(ref_loss.detach())) / max(1e-12, abs(float(ref_loss.detach())))
    worst, fails = (0.0, None), []
    for k, p in algo.unet.named_parameters():
        gr, gg = ref_sd[k].grad.float().flatten(), p.grad.detach().float().cpu().flatten()
        nr = gr.norm().item()
        if nr < 1e-12:
            assert gg.norm().item() < 1e-6, k
            continue
        e = (gg - gr).norm().item() / nr
        c = torch.dot(gg, gr).item() / (gg.norm().item() * nr + 1e-30)
        if e > worst[0]:
            worst = (e, k)
        if e > 6e-2 or c < 0.998:
            fails.append((k, round(e, 4), round(c, 5)))
    _record("train_368x768", {"loss": float(loss.detach()), "loss_ref": float(ref_loss.detach()), "loss_rel_err": rel,
                              "worst_grad_rel_l2": worst[0], "worst_grad_tensor": worst[1]})
    assert rel <= 2e-3, (float(loss.detach()), float(ref_loss.detach()))
    assert not fails, f"{len(fails)} parameter gradients out of tolerance (worst {worst}): {fails[:12]}"


def test_config4_full_size_vs_oracle():
    """(d).  BASELINE configs[3] inputs (flow ~ N(0, 4^2) px seed 3, frames ~ U[0,1], flow_gt = flow +
    N(0,1)) at 8x436x1024: warped image and mask bit-exact; loss values to 1e-5 relative; gradients wrt flow and frame2
    to 2e-4 of their max (fp32 atomics reorder the sums; the bound of tests/test_gpu_warp.py)."""
    from opticalflowdiffusion_b200 import warp as Wp
    _all_threads()
    B, H, W = 8, 436, 1024
    g = torch.Generator().manual_seed(3)
    flow = torch.randn(B, 2, H, W, generator=g) * 4
    f1 = torch.rand(B, 3, H, W, generator=g)
    f2 = torch.rand(B, 3, H, W, generator=g)
    gt = flow + torch.randn(B, 2, H, W, generator=g)
    ro, rm = O.backwarp(f2, flow)
    out, mask = Wp.warp_backward_flow(None, f2.cuda(), flow.cuda())
    assert torch.equal(out.cpu(), ro), "backward warp differs from the reference op sequence at 8x436x1024"
    assert torch.equal(mask.cpu(), rm)
    oob = 1.0 - rm.mean().item()
    fr, f2r = flow.clone().requires_grad_(True), f2.clone().requires_grad_(True)
    p_ref, e_ref, _, _ = O.photometric_epe(f1, f2r, fr, gt)
    (p_ref + e_ref).backward()
    fd, f2d = flow.cuda().requires_grad_(True), f2.cuda().requires_grad_(True)
    p, e = Wp.photometric_epe(f1.cuda(), f2d, fd, gt.cuda())
    (p + e).backward()
    np.testing.assert_allclose(p.item(), p_ref.item(), rtol=1e-5)
    np.testing.assert_allclose(e.item(), e_ref.item(), rtol=1e-5)
    gf_err = (fd.grad.cpu() - fr.grad).abs().max().item() / fr.grad.abs().max().item()
    g2_err = (f2d.grad.cpu() - f2r.grad).abs().max().item() / f2r.grad.abs().max().item()
    _record("config4_8x436x1024", {"backwarp": "bit-exact", "masked_out_fraction": oob, "photo": p.item(), "photo_ref": p_ref.item(),
                                   "epe": e.item(), "epe_ref": e_ref.item(), "gflow_max_err_over_max": gf_err,
                                   "gframe2_max_err_over_max": g2_err})
    assert gf_err <= 2e-4 and g2_err <= 2e-4, (gf_err, g2_err)


def test_config5_256x512_end_to_end_and_full_size_attention():
    """BASELINE configs[4] (matrix_flow resolution, 1024x2048).  The reference cannot run that size (its N x N attention
    matrix is 17 GB per sample), so SURVEY.md 8d row #5 asks for end-to-end parity at 256x512 and per-layer parity at
    full size:
      * DDIM-10 at 256x512 against the oracle trajectory: EPE <= 0.25 px, teacher-forced first step <= 3e-2;
      * the mid attention core at its full-size token count N = 32768 (128x256 tokens, 4 heads x 32) against an exact fp32
        softmax attention computed in query chunks on the GPU: relative L2 error <= 2e-2 (the bound of the small-shape
        tests in tests/test_gpu_unet_ops.py)."""
    from opticalflowdiffusion_b200 import _lib as L
    _all_threads()
    H, W, S = 256, 512, 10
    algo, sd = _build(2, S)
    sched = O.make_schedule(1000)
    cond = O.synthetic_frames(1, H, W, seed=300) * 2 - 1
    x_T = torch.randn(1, 2, H, W, generator=torch.Generator().manual_seed(77))
    with torch.no_grad():
        ref_traj, _ = O.ddim_sample(sd, sched, x_T, cond, 1000, S, return_all=True)
        t0 = O.ddim_times(1000, S)[0]
        ref0 = O.unet_forward(sd, x_T, cond, torch.full((1,), t0, dtype=torch.long))
        out0 = algo.unet(x_T.cuda(), cond.cuda(), torch.full((1,), t0, device="cuda", dtype=torch.long)).cpu()
    tf_err = (out0 - ref0).abs().max().item()
    out = algo.model.sample(1, external_cond=cond.cuda(), x_T=x_T)
    epe = _epe(out.cpu(), ref_traj[:, -1])
    # full-size attention layer
    lib = L.load()
    Ht, Wt = 128, 256
    g = torch.Generator().manual_seed(9)
    qkv = (torch.randn(1, Ht * Wt, 384, generator=g) * 1.5).to(torch.bfloat16).cuda()
    att = torch.empty(1, Ht * Wt, 128, device="cuda", dtype=torch.bfloat16)
    L.check(lib.fd_attention(L.ptr(qkv), L.ptr(att), 1, Ht * Wt, L.stream()))
    q, k, v = (t.float().reshape(Ht * Wt, 4, 32).permute(1, 0, 2) for t in qkv[0].split(128, dim=-1))     # (4, N, 32)
    ref = torch.empty(4, Ht * Wt, 32, device="cuda")
    for lo in range(0, Ht * Wt, 2048):
        sim = torch.einsum("hid,hjd->hij", q[:, lo:lo + 2048] * 32 ** -0.5, k)
        ref[:, lo:lo + 2048] = torch.einsum("hij,hjd->hid", sim.softmax(dim=-1), v)
    ref = ref.permute(1, 0, 2).reshape(Ht * Wt, 128)
    rel = ((att[0].float() - ref).norm() / ref.norm()).item()
    _record("config5", {"ddim10_256x512_epe_px": epe, "teacher_forced_256x512_max_abs_err": tf_err,
                        "attention_N32768_rel_l2": rel})
    assert tf_err <= 3e-2, tf_err
    assert epe <= 0.25, epe
    assert rel <= 2e-2, rel
