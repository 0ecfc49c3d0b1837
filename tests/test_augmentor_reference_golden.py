"""The augmentation against goldens produced by the reference's OWN ``Augmentor``
(oracle/make_goldens_augment.py; augmentation.py:6-76): the per-item torchvision restatement ``flow_diffuser.Augmentor``
(CPU, bit-exact: same torchvision calls, same RNG consumption) and, on the GPU, ``augment.GpuAugmentor`` (<= 3e-5)."""
import random

import numpy as np
import pytest
import torch


def batch(seed, B, S):
    g = torch.Generator().manual_seed(1000 + seed)
    return (torch.rand(B, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g), torch.randn(B, 2, S, S, generator=g) * 4)


def test_host_augmentor_equals_the_reference(golden):
    from opticalflowdiffusion_b200.flow_diffuser import Augmentor
    g = golden("augmentor_ref")
    B, S = int(g["B"]), int(g["S"])
    for seed in g["seeds"]:
        seed = int(seed)
        random.seed(seed)
        torch.manual_seed(seed)
        aug = Augmentor()
        out = aug(tuple(t.clone() for t in batch(seed, B, S)))
        for name, o in zip(("img", "tgt", "flow"), out):
            np.testing.assert_array_equal(o.numpy(), g[f"{name}_{seed}"], err_msg=f"{name} seed {seed}")


@pytest.mark.gpu
def test_gpu_augmentor_equals_the_reference(golden):
    from opticalflowdiffusion_b200.augment import GpuAugmentor
    g = golden("augmentor_ref")
    B, S = int(g["B"]), int(g["S"])
    worst = 0.0
    for seed in g["seeds"]:
        seed = int(seed)
        random.seed(seed)
        torch.manual_seed(seed)
        aug = GpuAugmentor()
        out = aug(tuple(t.cuda() for t in batch(seed, B, S)))
        for name, o in zip(("img", "tgt", "flow"), out):
            err = np.abs(o.cpu().numpy() - g[f"{name}_{seed}"]).max()
            worst = max(worst, float(err))
            assert err <= 3e-5, (name, seed, err)
    print("GpuAugmentor vs the reference Augmentor: worst |err|", worst)
