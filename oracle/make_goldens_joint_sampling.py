"""TEST INFRASTRUCTURE ONLY.  Golden trajectory of the reference's sampler with ``target: joint`` (the reference's default,
configurations/algorithm/flow_diffuser.yaml:15): ``ConditionalDiffusion.p_sample_loop`` (denoising_diffusion.py:700-729) around
``UnetWithWarp`` (flow_diffuser.py:20-63) -- the state carries the NaN holes of the forward splat from step to step, the UNet
sees NaN -> 0 plus the mask channel, and the flow channels of the state are the prediction the next step starts from.

As shipped the reference samples with DDPM over all ``timesteps`` (sampling_timesteps is never passed, flow_diffuser.py:116-126);
a model with timesteps = 5 keeps the golden small.  Noise draws are fed through ``patched_randn`` so that they can be replayed;
the forward splat is the reference's own kernels compiled for the host (oracle/build_ref.py).

Run in the build container:  python oracle/make_goldens_joint_sampling.py  ->  tests/golden/joint_ddpm5_32x48.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_stubs  # noqa: E402
from oracle.make_goldens import patched_randn, quiet, weight_checksums  # noqa: E402
from oracle.make_goldens_flow_learner import HostSplat  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
HEAD_SCALE = 1.5


def main():
    ns = ref_stubs.import_reference()
    build_ref.build()
    ns.softsplat_new.softsplat_func = HostSplat
    B, H, W, T = 2, 32, 48, 5
    torch.manual_seed(0)
    cfg = ref_stubs.reference_cfg(target="joint", image_size=64, zero_init=False, timesteps=T)
    m = ns.flow_diffuser.FlowDiffuser(cfg)
    # a random-init UNet predicts flows of ~0.1 (x flow_max = 2 px): scale the head so that the splat makes real holes
    with torch.no_grad():
        m.unet.final_conv.weight.mul_(HEAD_SCALE)
    sums, asums = weight_checksums(m.unet.state_dict())
    g = torch.Generator().manual_seed(77)
    cond = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    x_T = torch.randn(B, 5, H, W, generator=g)
    noises = [torch.randn(B, 5, H, W, generator=g) for _ in range(T - 1)]
    queue = [x_T] + list(noises)
    with torch.no_grad(), patched_randn(queue):
        traj = quiet(m.model.p_sample_loop, (B, 5, H, W), return_all_timesteps=True, external_cond=cond)
    assert not queue, len(queue)
    traj = traj.numpy()
    print("traj", traj.shape, "NaN fraction per step", [float(np.isnan(traj[:, i]).mean()) for i in range(traj.shape[1])],
          "flow range", float(np.nanmin(traj[:, -1, 3:])), float(np.nanmax(traj[:, -1, 3:])))
    np.savez_compressed(os.path.join(GOLD, "joint_ddpm5_32x48.npz"), seed=0, head_scale=HEAD_SCALE, w_sums=sums, w_asums=asums,
                        cond=cond.numpy(), x_T=x_T.numpy(), noises=torch.stack(noises).numpy(), traj=traj)


if __name__ == "__main__":
    main()
