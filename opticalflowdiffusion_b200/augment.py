"""GPU-side training augmentation.

``GpuAugmentor`` is the reference's ``Augmentor`` (algorithms/diffusion_animation/augmentation.py:6-76) with the image
arithmetic moved into four launches of ``libflowdiff.so`` (``csrc/fd_augment.cu``) over the whole batch, instead of
~30 torchvision calls per item from Python.  What stays on the host is only the *sampling of parameters*, done with
the very samplers the reference's transforms use (``ColorJitter.get_params``, ``GaussianBlur.get_params``,
``RandomResizedCrop.get_params``, ``torch.rand(1) < p``) in the same order, so under a given ``random`` / ``torch``
seed it takes exactly the decisions of the torchvision path (``flow_diffuser.Augmentor``) -- which is how
``tests/test_gpu_augment.py`` checks it.
"""
from __future__ import annotations

import random
from typing import List, Tuple

import torch

from . import _lib

Tensor = torch.Tensor


class GpuAugmentor:
    def __init__(self):
        import torchvision.transforms as T
        self._T = T
        lim = 0.1
        deltas = [(random.random() - 0.5) * 2 * lim for _ in range(4)]          # augmentation.py:14-19
        ranges = [(b + d, b + d + 0.01) for b, d in zip((1, 1, 1, 0), deltas)]
        jit = T.ColorJitter(*ranges)                                           # validates / normalises the ranges
        self._jitter_ranges = (jit.brightness, jit.contrast, jit.saturation, jit.hue)
        self._sigma = random.random() * 0.5                                    # augmentation.py:26
        T.GaussianBlur(3, self._sigma)                                         # same validation as the reference
        self.last_plan = None

    # ------------------------------------------------------------------ host: decisions only
    def plan(self, B: int, H: int, W: int) -> Tuple[List[int], List[float], List[int]]:
        T = self._T
        fints = [[0, 0, 1, 2, 3, 0, 0, 0] for _ in range(2 * B)]
        ffloats = [[1.0, 1.0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0] for _ in range(2 * B)]
        iints = [[0] * 8 for _ in range(B)]
        for i in range(B):                                  # image_augs, item by item (augmentation.py:29-33,57-66)
            if torch.rand(1) < 0.4:
                for k in range(2):                          # each frame of the pair draws its own order and factors
                    fn_idx, b, c, s, h = T.ColorJitter.get_params(*self._jitter_ranges)
                    f = 2 * i + k
                    fints[f][0] = 1
                    fints[f][1:5] = [int(v) for v in fn_idx]
                    ffloats[f][0:4] = [float(b), float(c), float(s), float(h)]
            if torch.rand(1) < 0.1:
                fints[2 * i][5] = fints[2 * i + 1][5] = 1
            if torch.rand(1) < 0.2:
                for k in range(2):
                    sigma = T.GaussianBlur.get_params(self._sigma, self._sigma)
                    x = torch.linspace(-1.0, 1.0, steps=3)
                    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
                    k1 = pdf / pdf.sum()
                    f = 2 * i + k
                    fints[f][6] = 1
                    ffloats[f][4], ffloats[f][5] = float(k1[0]), float(k1[1])
        probe = torch.empty((1, 1, H, W))
        for i in range(B):                                  # whole_augs (augmentation.py:34-56)
            # RandomApply draws once; the RandomHorizontalFlip(p=1.0) / RandomVerticalFlip(p=1.0) it wraps draws AGAIN
            # (augmentation.py:34-42): the second number is consumed to keep the stream aligned with the reference
            # (found by tests/test_augmentor_reference_golden.py against the reference's own Augmentor)
            iints[i][0] = int(torch.rand(1) < 0.3)
            if iints[i][0]:
                torch.rand(1)
            iints[i][1] = int(torch.rand(1) < 0.3)
            if iints[i][1]:
                torch.rand(1)
            if torch.rand(1) < 0.15:
                top, left, h, w = T.RandomResizedCrop.get_params(probe, [0.8, 1.0], [0.9, 1.1])
                iints[i][2:7] = [1, int(top), int(left), int(h), int(w)]
        return ([v for row in fints for v in row], [v for row in ffloats for v in row], [v for row in iints for v in row])

    # ------------------------------------------------------------------ device
    def __call__(self, batch):
        img, tgt, flow = batch
        _lib.require_cuda(img, tgt, flow)
        lib = _lib.load(check_device=True)
        B, C, H, W = img.shape
        assert C == 3 and tgt.shape == img.shape and flow.shape == (B, 2, H, W)
        dev = img.device
        img, tgt, flow = (t.detach().float().contiguous() for t in (img, tgt, flow))
        fints, ffloats, iints = self.plan(B, H, W)
        self.last_plan = (fints, ffloats, iints)
        ints = torch.tensor(fints + iints, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
        floats = torch.tensor(ffloats, dtype=torch.float32).pin_memory().to(dev, non_blocking=True)
        fi_ptr, ii_ptr = ints.data_ptr(), ints.data_ptr() + 4 * len(fints)
        st = _lib.stream()
        frames = torch.empty(B, 2, 3, H, W, device=dev, dtype=torch.float32)
        means = torch.empty(2 * B, device=dev, dtype=torch.float32)
        _lib.check(lib.fd_aug_photometric(_lib.ptr(img), _lib.ptr(tgt), fi_ptr, _lib.ptr(floats), _lib.ptr(means),
                                          _lib.ptr(frames), B, H * W, st))
        blurred = frames
        if any(fints[8 * f + 6] for f in range(2 * B)):
            blurred = torch.empty_like(frames)
            _lib.check(lib.fd_aug_blur3(_lib.ptr(frames), fi_ptr, _lib.ptr(floats), _lib.ptr(blurred), B, H, W, st))
        o_img, o_tgt, o_flow = torch.empty_like(img), torch.empty_like(tgt), torch.empty_like(flow)
        _lib.check(lib.fd_aug_geometric(_lib.ptr(frames), _lib.ptr(blurred), _lib.ptr(flow), fi_ptr, ii_ptr, _lib.ptr(o_img),
                                        _lib.ptr(o_tgt), _lib.ptr(o_flow), B, H, W, st))
        return o_img, o_tgt, o_flow
