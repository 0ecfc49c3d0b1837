"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the joint / target pyramid loss of flow_diffuser
(``ConditionalDiffusion.p_losses`` -> ``_loss``, denoising_diffusion.py:823-983, with ``UnetWithWarp``,
flow_diffuser.py:20-63) from the UNMODIFIED reference classes.  ``target: joint`` is the reference's default
(configurations/algorithm/flow_diffuser.yaml:15).

The reference's forward splat is CuPy/CUDA-only (softsplat_new.py:444-445), so ``softsplat_new.softsplat_func`` is swapped
for an autograd Function over the reference's own three kernels compiled for the host (oracle/build_ref.py), exactly as
oracle/make_goldens_flow_learner.py does; everything else (FlowDiffuser.preprocess, q_sample, UnetWithWarp, the level loop,
nan_mse, nanmean) runs as shipped.  ``p_losses`` is called directly, which bypasses only the square-image assert of
``forward`` (:986-987).

Run in the build container:  python oracle/make_goldens_joint.py  ->  tests/golden/p_losses_{joint,target}_32x48.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_stubs  # noqa: E402
from oracle.make_goldens import quiet, weight_checksums  # noqa: E402
from oracle.make_goldens_flow_learner import HostSplat  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def one(ns, target: str):
    B, H, W = 2, 32, 48
    torch.manual_seed(0)
    cfg = ref_stubs.reference_cfg(target=target, image_size=64, zero_init=False)
    m = ns.flow_diffuser.FlowDiffuser(cfg)
    sums, asums = weight_checksums(m.unet.state_dict())
    g = torch.Generator().manual_seed(41)
    img = torch.rand(B, 3, H, W, generator=g)
    tgt = torch.rand(B, 3, H, W, generator=g)
    flow = torch.randn(B, 2, H, W, generator=g) * 3
    t = torch.tensor([620, 85], dtype=torch.long)
    first, cond, flow_n = quiet(m.preprocess, (img, tgt, flow), aug=False)       # flow_diffuser.py:136-168
    noise = torch.randn(first.shape, generator=g)
    # capture the UNet's raw flow prediction and its gradient
    grabbed = {}

    def hook(_m, _i, out):
        out.retain_grad()
        grabbed["flow_pred"] = out
    h = m.unet.register_forward_hook(hook)
    m.zero_grad()
    kw = dict(additional_tgt=flow_n, additional_weight=cfg.flow_weight) if target == "target" else {}
    loss = quiet(m.model.p_losses, first, t, noise=noise.clone(), external_cond=cond, **kw)
    quiet(loss.backward)
    h.remove()
    torch.autograd.set_detect_anomaly(False)
    fp = grabbed["flow_pred"]
    # the loss with the model output overridden by the ground truth ("val/ideal_loss", flow_diffuser.py:256-259)
    with torch.no_grad():
        warped = m._model._warp(cond, flow_n)
        override = (warped, flow_n) if target == "target" else (torch.cat((warped, flow_n), dim=1), None)
        ideal = quiet(m.model.p_losses, first, t, noise=noise.clone(), external_cond=cond, model_out_override=override, **kw)
    torch.autograd.set_detect_anomaly(False)
    unet = m.unet
    d = dict(seed=0, w_sums=sums, w_asums=asums, img=img.numpy(), tgt=tgt.numpy(), flow=flow.numpy(), t=t.numpy(),
             noise=noise.numpy(), first=first.detach().numpy(), cond=cond.numpy(), flow_n=flow_n.numpy(),
             flow_pred=fp.detach().numpy(), grad_flow_pred=fp.grad.numpy(), loss=np.array(float(loss)),
             ideal_loss=np.array(float(ideal)),
             grad_final_conv_w=unet.final_conv.weight.grad.numpy(), grad_final_conv_b=unet.final_conv.bias.grad.numpy(),
             grad_init_conv_b=unet.init_conv.bias.grad.numpy(),
             grad_mid_qkv_w_sum=np.array(float(unet.mid_attn.fn.fn.to_qkv.weight.grad.double().abs().sum())))
    name = f"p_losses_{target}_32x48.npz"
    np.savez_compressed(os.path.join(GOLD, name), **d)
    print(name, "loss", float(loss), "ideal", float(ideal), "nan fraction of x0", float(torch.isnan(first).float().mean()),
          "|grad flow_pred| max", float(fp.grad.abs().max()))


def main():
    ns = ref_stubs.import_reference()
    build_ref.build()
    ns.softsplat_new.softsplat_func = HostSplat
    for target in ("joint", "target"):
        one(ns, target)


if __name__ == "__main__":
    main()
