"""TEST INFRASTRUCTURE ONLY.  Golden vectors for FlowLearner.loss (flow_learner.py:133-222) from the
UNMODIFIED reference class: the reference's softsplat is CuPy-only, so ``softsplat_new.softsplat_func`` is swapped for an
autograd Function over the reference's own three kernels compiled for the host (oracle/build_ref.py); everything else
(Unet, UnetWithWarp, the 832-term loop, fill_holes_nan, nan_charbonnier, edgeaware_smoothness1) runs as shipped.

Run in the build container:  python oracle/make_goldens_flow_learner.py  ->  tests/golden/flow_learner_32x32.npz
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_stubs  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class HostSplat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tenIn, tenFlow, scale, offset_x, offset_y):
        tenIn, tenFlow = tenIn.float().contiguous(), tenFlow.float().contiguous()
        ctx.save_for_backward(tenIn, tenFlow)
        ctx.geom = (int(scale), int(offset_x), int(offset_y))
        return build_ref.ref_splat_out(tenIn, tenFlow, *ctx.geom)

    @staticmethod
    def backward(ctx, gout):
        tenIn, tenFlow = ctx.saved_tensors
        gin, gflow = build_ref.ref_splat_backward(tenIn, tenFlow, gout.float().contiguous(), *ctx.geom)
        return gin, gflow, None, None, None


class Cfg:
    flow_max = 20
    latent = False
    zero_init = False
    c2f = False
    lr = 8e-5
    weight_decay = 1e-6
    sparsity_weight = 0.0
    occlusion_mask = True
    train_aug = False
    image_size = 32


def main():
    import importlib
    ns = ref_stubs.import_reference()
    build_ref.build()
    ns.softsplat_new.softsplat_func = HostSplat
    fl = importlib.import_module("algorithms.diffusion_animation.flow_learner")
    random.seed(0)
    torch.manual_seed(0)
    m = fl.FlowLearner(Cfg())
    sd = m.unet.model.state_dict()
    sums = np.array([float(v.double().sum()) for v in sd.values()])
    g = torch.Generator().manual_seed(21)
    B, H, W = 1, 32, 32
    img = torch.rand(B, 3, H, W, generator=g)
    tgt = torch.rand(B, 3, H, W, generator=g)
    flow = torch.randn(B, 2, H, W, generator=g) * 3
    tgt_, cond, flow_ = m.preprocess((img, tgt, flow), aug=False)
    with torch.no_grad():
        out = m.model(cond, additional_out=True)
    m.zero_grad()
    loss = m.loss(tgt_, cond, flow_)
    loss.backward()
    torch.autograd.set_detect_anomaly(False)
    unet = m.unet.model
    d = dict(seed=0, w_sums=sums, img=img.numpy(), tgt=tgt.numpy(), flow=flow.numpy(), model_out=out.numpy(),
             loss=np.array(float(loss)), ideal_loss=np.array(float(m.loss(tgt_, cond, flow_, override_flow=flow_).detach())),
             grad_final_conv_w=unet.final_conv.weight.grad.numpy(), grad_final_conv_b=unet.final_conv.bias.grad.numpy(),
             grad_init_conv_b=unet.init_conv.bias.grad.numpy(),
             grad_mid_qkv_w_sum=np.array(float(unet.mid_attn.fn.fn.to_qkv.weight.grad.double().abs().sum())),
             n_keys=np.array(len(m.state_dict())))
    np.savez_compressed(os.path.join(GOLD, "flow_learner_32x32.npz"), **d)
    print("flow_learner_32x32.npz loss", float(loss), "ideal", float(d["ideal_loss"]), "keys", len(m.state_dict()))


if __name__ == "__main__":
    main()
