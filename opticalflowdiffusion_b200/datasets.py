"""Synthetic Sintel-shaped data: smooth random frame pairs with a known flow.
Replaces the reference's datasets/, whose Sintel loader reads hard-coded local paths
(datasets/animation/sintel.py:19-21,73).  Items are ``(img, tgt, flow)`` like FlyingChairs/KITTI:
``img, tgt`` in [0,1] (3,H,W), ``flow`` (2,H,W) in pixels."""
from __future__ import annotations

import math

import torch


def synthetic_frames(batch: int, height: int, width: int, seed: int = 0) -> torch.Tensor:
    """Sum of 4 low-frequency sinusoids per channel + U[0,0.05] noise, clamped to [0,1]."""
    g = torch.Generator().manual_seed(seed)
    yy = torch.linspace(0, 1, height).view(1, 1, height, 1)
    xx = torch.linspace(0, 1, width).view(1, 1, 1, width)
    img = torch.zeros(batch, 3, height, width)
    for _ in range(4):
        fx = torch.rand(batch, 3, 1, 1, generator=g) * 6.0
        fy = torch.rand(batch, 3, 1, 1, generator=g) * 6.0
        ph = torch.rand(batch, 3, 1, 1, generator=g) * (2 * math.pi)
        img = img + 0.125 * torch.sin(2 * math.pi * (fx * xx + fy * yy) + ph)
    img = img + 0.5 + torch.rand(batch, 3, height, width, generator=g) * 0.05
    return img.clamp(0.0, 1.0)


class SyntheticSintelDataset(torch.utils.data.Dataset):
    def __init__(self, cfg, split: str = "training", device=None):
        self.h, self.w = int(cfg.height), int(cfg.width)
        self.n = int(cfg.length)
        self.sigma = float(cfg.flow_sigma)
        self.seed = int(cfg.seed) + {"training": 0, "validation": 1, "test": 2}.get(split, 3)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        img = synthetic_frames(1, self.h, self.w, seed=self.seed * 100003 + i)[0]
        # smooth flow: coarse noise upsampled bilinearly
        coarse = torch.randn(1, 2, max(2, self.h // 32), max(2, self.w // 32), generator=g) * self.sigma
        flow = torch.nn.functional.interpolate(coarse, size=(self.h, self.w), mode="bilinear", align_corners=True)[0]
        tgt = synthetic_frames(1, self.h, self.w, seed=self.seed * 100003 + i + 50021)[0]
        return img, tgt, flow
