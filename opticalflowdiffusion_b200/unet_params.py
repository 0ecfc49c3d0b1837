"""Parameter tree of the noise-prediction UNet.

Holds the 276 fp32 master tensors under exactly the names the reference's ``Unet`` registers
(denoising_diffusion.py:272-361: ``init_conv.weight`` ... ``downs.0.2.fn.fn.to_out.1.g`` ...
``final_conv.bias``) so reference checkpoints load with ``load_state_dict`` and so that
``torch.manual_seed(s)`` followed by construction yields the reference's random init: layers
are created in the reference's order with the same ``nn.Conv2d`` / ``nn.Linear`` initialisers.

The modules here are *containers only* -- none of them has a ``forward``; the arithmetic is in
``unet.py`` which walks this tree and launches the CUDA kernels.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
from torch import nn


class _Holder(nn.Module):
    """A module that only owns parameters / children."""

    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("parameter container; use opticalflowdiffusion_b200.unet.Unet")


class _Gain(_Holder):
    """Channel LayerNorm gain ``g`` of shape (1, C, 1, 1) (denoising_diffusion.py:116-119)."""

    def __init__(self, dim: int):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))


class _Block(_Holder):
    """proj (weight-standardised 3x3) + GroupNorm(8) affine (denoising_diffusion.py:172-177)."""

    def __init__(self, cin: int, cout: int, groups: int = 8):
        super().__init__()
        self.proj = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm = nn.GroupNorm(groups, cout)


class _ResnetBlock(_Holder):
    """mlp.1 / block1 / block2 / res_conv (denoising_diffusion.py:190-200)."""

    def __init__(self, cin: int, cout: int, time_dim, groups: int = 8):
        super().__init__()
        # time_emb_dim = None (Unet(time_in=False), :193-196) -> no mlp, Block1 runs without (scale, shift)
        self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, cout * 2)) if time_dim is not None else None
        self.block1 = _Block(cin, cout, groups)
        self.block2 = _Block(cout, cout, groups)
        self.res_conv = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()
        self.cin, self.cout = cin, cout


class _LinearAttention(_Holder):
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32):
        super().__init__()
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), _Gain(dim))
        self.heads, self.dim_head, self.dim = heads, dim_head, dim


class _Attention(_Holder):
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32):
        super().__init__()
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)
        self.heads, self.dim_head, self.dim = heads, dim_head, dim


class _PreNorm(_Holder):
    def __init__(self, dim: int, fn: nn.Module):
        super().__init__()
        self.fn = fn
        self.norm = _Gain(dim)


class _Residual(_Holder):
    def __init__(self, fn: nn.Module):
        super().__init__()
        self.fn = fn


def _resample(conv: nn.Conv2d) -> nn.Sequential:
    """``Sequential(<rearrange|upsample>, Conv2d)``: the conv sits at child '1'
    (denoising_diffusion.py:89-99)."""
    return nn.Sequential(nn.Identity(), conv)


class UnetParams(_Holder):
    """dim=64, dim_mults=(1,2,4,8), resnet groups 8, sinusoidal time embedding."""

    def __init__(self, dim: int = 64, channels: int = 5, out_dim: int = 2,
                 dim_mults: Tuple[int, ...] = (1, 2, 4, 8), groups: int = 8, time_in: bool = True):
        super().__init__()
        self.dim, self.channels, self.out_dim, self.time_in = dim, channels, out_dim, time_in
        self.init_conv = nn.Conv2d(channels, dim, 7, padding=3)
        dims: List[int] = [dim] + [dim * m for m in dim_mults]
        self.in_out = list(zip(dims[:-1], dims[1:]))
        time_dim = dim * 4 if time_in else None          # :306-318: no time_mlp at all when time_in is False
        self.time_dim = time_dim
        if time_in:
            self.time_mlp = nn.Sequential(nn.Identity(), nn.Linear(dim, time_dim), nn.GELU(),
                                          nn.Linear(time_dim, time_dim))
        self.downs = nn.ModuleList()
        self.ups = nn.ModuleList()
        n = len(self.in_out)
        for i, (cin, cout) in enumerate(self.in_out):
            last = i >= n - 1
            self.downs.append(nn.ModuleList([
                _ResnetBlock(cin, cin, time_dim, groups),
                _ResnetBlock(cin, cin, time_dim, groups),
                _Residual(_PreNorm(cin, _LinearAttention(cin))),
                _resample(nn.Conv2d(cin * 4, cout, 1)) if not last else nn.Conv2d(cin, cout, 3, padding=1),
            ]))
        mid = dims[-1]
        self.mid_block1 = _ResnetBlock(mid, mid, time_dim, groups)
        self.mid_attn = _Residual(_PreNorm(mid, _Attention(mid)))
        self.mid_block2 = _ResnetBlock(mid, mid, time_dim, groups)
        for i, (cin, cout) in enumerate(reversed(self.in_out)):
            last = i == n - 1
            self.ups.append(nn.ModuleList([
                _ResnetBlock(cout + cin, cout, time_dim, groups),
                _ResnetBlock(cout + cin, cout, time_dim, groups),
                _Residual(_PreNorm(cout, _LinearAttention(cout))),
                _resample(nn.Conv2d(cout, cin, 3, padding=1)) if not last else nn.Conv2d(cout, cin, 3, padding=1),
            ]))
        self.final_res_block = _ResnetBlock(dim * 2, dim, time_dim, groups)
        self.final_conv = nn.Conv2d(dim, out_dim, 1)
