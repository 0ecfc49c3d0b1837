"""-m gpu: the fused GPU augmentation (augment.GpuAugmentor / csrc/fd_augment.cu) against the
torchvision calls the reference's Augmentor makes (augmentation.py:6-76).

Tolerance: fp32 image arithmetic re-associated inside one kernel (FMA contraction) -> max |err| <= 2e-5 on [0,1]
images / pixel-unit flows; flips and the no-op path are bit-exact."""
import random

import pytest
import torch
import torchvision.transforms.functional as TF

pytestmark = pytest.mark.gpu


def data(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, H, W, generator=g).cuda()
    tgt = torch.rand(B, 3, H, W, generator=g).cuda()
    flow = (torch.randn(B, 2, H, W, generator=g) * 4).cuda()
    return img, tgt, flow


def run_tables(img, tgt, flow, fints, ffloats, iints):
    from opticalflowdiffusion_b200 import _lib
    lib = _lib.load(check_device=True)
    B, _, H, W = img.shape
    fi = torch.tensor(fints, dtype=torch.int32).cuda()
    ff = torch.tensor(ffloats, dtype=torch.float32).cuda()
    ii = torch.tensor(iints, dtype=torch.int32).cuda()
    frames = torch.empty(B, 2, 3, H, W, device="cuda")
    blurred = torch.empty_like(frames)
    means = torch.empty(2 * B, device="cuda")
    st = _lib.stream()
    _lib.check(lib.fd_aug_photometric(_lib.ptr(img), _lib.ptr(tgt), _lib.ptr(fi), _lib.ptr(ff), _lib.ptr(means), _lib.ptr(frames),
                                      B, H * W, st))
    _lib.check(lib.fd_aug_blur3(_lib.ptr(frames), _lib.ptr(fi), _lib.ptr(ff), _lib.ptr(blurred), B, H, W, st))
    o = [torch.empty_like(t) for t in (img, tgt, flow)]
    _lib.check(lib.fd_aug_geometric(_lib.ptr(frames), _lib.ptr(blurred), _lib.ptr(flow), _lib.ptr(fi), _lib.ptr(ii),
                                    _lib.ptr(o[0]), _lib.ptr(o[1]), _lib.ptr(o[2]), B, H, W, st))
    torch.cuda.synchronize()
    return frames, o


def test_photometric_ops_vs_torchvision():
    B, H, W = 3, 21, 34
    img, tgt, flow = data(B, H, W, 0)
    orders = [[0, 1, 2, 3], [3, 1, 0, 2], [1, 3, 2, 0], [2, 0, 3, 1], [3, 2, 1, 0], [1, 0, 2, 3]]
    factors = [[1.07, 0.93, 1.05, 0.04], [0.91, 1.1, 0.95, -0.07], [1.0, 1.02, 1.09, 0.1], [1.1, 0.9, 0.9, -0.1],
               [0.95, 1.05, 1.0, 0.02], [1.03, 0.97, 1.08, -0.03]]
    fints, ffloats = [], []
    for f in range(2 * B):
        fints.append([1] + orders[f] + [int(f == 3), 0, 0])
        ffloats.append(factors[f] + [0.0, 1.0, 0.0, 0.0])
    iints = [[0] * 8 for _ in range(B)]
    frames, (o_img, o_tgt, o_flow) = run_tables(img, tgt, flow, fints, ffloats, iints)
    ops = [TF.adjust_brightness, TF.adjust_contrast, TF.adjust_saturation, TF.adjust_hue]
    for f in range(2 * B):
        x = (tgt if f & 1 else img)[f // 2:f // 2 + 1]
        for k in orders[f]:
            x = ops[k](x, factors[f][k])
        if f == 3:
            x = TF.rgb_to_grayscale(x, 3)
        err = (frames[f // 2, f & 1] - x[0]).abs().max().item()
        assert err <= 2e-5, (f, err)
    assert torch.equal(o_img, frames[:, 0]) and torch.equal(o_tgt, frames[:, 1]) and torch.equal(o_flow, flow)


def test_blur_flips_crop_vs_torchvision():
    B, H, W = 4, 26, 39
    img, tgt, flow = data(B, H, W, 1)
    sigmas = [0.37, 0.05, 0.49]
    fints = [[0, 0, 1, 2, 3, 0, 0, 0] for _ in range(2 * B)]
    ffloats = [[1, 1, 1, 0, 0, 1, 0, 0] for _ in range(2 * B)]
    for f, sg in zip((0, 1, 5), sigmas):
        x = torch.linspace(-1.0, 1.0, steps=3)
        pdf = torch.exp(-0.5 * (x / sg).pow(2))
        k1 = pdf / pdf.sum()
        fints[f][6] = 1
        ffloats[f][4], ffloats[f][5] = float(k1[0]), float(k1[1])
    crops = {1: (3, 5, 20, 31), 3: (0, 0, 26, 35)}
    iints = [[1, 0, 0, 0, 0, 0, 0, 0], [0, 1, 1, 3, 5, 20, 31, 0], [1, 1, 0, 0, 0, 0, 0, 0], [1, 0, 1, 0, 0, 26, 35, 0]]
    _, (o_img, o_tgt, o_flow) = run_tables(img, tgt, flow, fints, ffloats, iints)
    for i in range(B):
        a, b = img[i:i + 1], tgt[i:i + 1]
        for f, sg in zip((0, 1, 5), sigmas):
            if f == 2 * i:
                a = TF.gaussian_blur(a, [3, 3], [sg, sg])
            if f == 2 * i + 1:
                b = TF.gaussian_blur(b, [3, 3], [sg, sg])
        item = torch.cat((a, b, flow[i:i + 1]), 1).clone()
        if iints[i][0]:
            item = TF.hflip(item)
            item[:, -1] = -item[:, -1]
        if iints[i][1]:
            item = TF.vflip(item)
            item[:, -2] = -item[:, -2]
        exact = True
        if iints[i][2]:
            top, left, h, w = crops[i]
            scale = torch.tensor([h / H, w / W], device="cuda")
            item[:, -2:] = item[:, -2:] * scale[None, :, None, None]
            item = TF.resized_crop(item, top, left, h, w, (H, W), antialias=False)
            exact = False
        got = torch.cat((o_img[i], o_tgt[i], o_flow[i]), 0)
        err = (got - item[0]).abs().max().item()
        if exact and not any(f // 2 == i for f in (0, 1, 5)):
            assert err == 0.0, (i, err)
        assert err <= 3e-5, (i, err)


@pytest.mark.parametrize("shape", [(4, 40, 56), (3, 32, 32)])
def test_gpu_augmentor_takes_the_decisions_of_the_torchvision_path(shape):
    """Same python / torch seeds -> the fused path and the per-item torchvision path (flow_diffuser.Augmentor, the
    restatement of augmentation.py) give the same batch; 24 seeds so that every branch is drawn."""
    from opticalflowdiffusion_b200.augment import GpuAugmentor
    from opticalflowdiffusion_b200.flow_diffuser import Augmentor
    B, H, W = shape
    seen = {"jitter": 0, "gray": 0, "blur": 0, "hflip": 0, "vflip": 0, "crop": 0}
    for seed in range(24):
        img, tgt, flow = data(B, H, W, 100 + seed)
        random.seed(seed)
        torch.manual_seed(seed)
        ref = Augmentor()((img.clone(), tgt.clone(), flow.clone()))
        state = torch.get_rng_state()
        random.seed(seed)
        torch.manual_seed(seed)
        aug = GpuAugmentor()
        got = aug((img, tgt, flow))
        assert torch.equal(torch.get_rng_state(), state)            # consumed exactly the same random numbers
        fints, _, iints = aug.last_plan
        for f in range(2 * B):
            seen["jitter"] += fints[8 * f]
            seen["gray"] += fints[8 * f + 5]
            seen["blur"] += fints[8 * f + 6]
        for i in range(B):
            seen["hflip"] += iints[8 * i]
            seen["vflip"] += iints[8 * i + 1]
            seen["crop"] += iints[8 * i + 2]
        for name, a, b in zip(("img", "tgt", "flow"), got, ref):
            err = (a - b).abs().max().item()
            assert err <= 3e-5, (seed, name, err)
    assert all(v > 0 for v in seen.values()), seen


def test_preprocess_uses_the_gpu_augmentor():
    from opticalflowdiffusion_b200 import FlowDiffuser
    from opticalflowdiffusion_b200.augment import GpuAugmentor
    from opticalflowdiffusion_b200.config import compose
    algo = FlowDiffuser(compose(["algorithm.target=flow"]).algorithm).cuda()
    assert isinstance(algo.augmentor, GpuAugmentor)
    img, tgt, flow = data(2, 32, 48, 7)
    first, cond, fl = algo.preprocess((img, tgt, flow), aug=True)
    assert first.shape == (2, 2, 32, 48) and cond.shape == (2, 3, 32, 48)
    assert float(fl.abs().max()) <= 1.0 and float(cond.min()) >= -1.0 and float(cond.max()) <= 1.0
