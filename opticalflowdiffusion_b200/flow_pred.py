"""Latent-mode autoencoder of the reference (``Autoencoder``, flow_pred.py:17-58): two three-level UNets
(``dim_mults=(1, 2, 4)``, no time input) on the same sm_100a kernels as the flow UNet.

* encoder ``Unet(64, channels=3, out_dim=latent_dim)``: image in [0, 1] -> latent clamped to [-1, 1];
* decoder ``Unet(64, channels=latent_dim + 3, out_dim=3)``: (latent, image) -> image in [0, 1]; its 19 input channels take
  the wide ``init_conv`` path (fd_pack_input_wide + 49-tap implicit GEMM);
* ``forward`` = encode, forward-splat the latent along the flow (warp.py 'forward' mode), decode.

Only what ``FlowDiffuser`` uses (flow_diffuser.py:81-95, 143-148, 255, 304-318) is provided: the autoencoder is a frozen
feature extractor there, so it always runs the inference kernels.  Training it (``FlowPred``, flow_pred.py:60-) is outside the
flow_diffuser path."""
from typing import Optional

import torch
from torch import Tensor, nn

from . import warp as W
from .unet import Unet


class Autoencoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.model_enc = Unet(64, channels=3, out_dim=cfg.latent_dim, dim_mults=(1, 2, 4), time_in=False)
        self.model_dec = Unet(64, channels=cfg.latent_dim + 3, dim_mults=(1, 2, 4), out_dim=3, time_in=False)

    @torch.no_grad()
    def encode(self, x: Tensor) -> Tensor:
        """flow_pred.py:49-50"""
        return torch.clamp(self.model_enc(2 * x - 1.0), -1.0, 1.0)

    @torch.no_grad()
    def decode(self, latent: Tensor, x: Tensor) -> Tensor:
        """flow_pred.py:53-57"""
        out = self.model_dec(torch.cat((latent, 2 * x - 1), dim=1).contiguous())
        return (torch.clamp(out, -1.0, 1.0) + 1.0) / 2.0

    @torch.no_grad()
    def forward(self, x: Tensor, flow: Tensor, return_latent: bool = False) -> Tensor:
        """flow_pred.py:38-47"""
        lat = W.warp(self.encode(x), None, flow, mode="forward")
        if return_latent:
            return lat
        return self.decode(lat, x)


def load_autoencoder(ae: Autoencoder, path: Optional[str]) -> bool:
    """flow_diffuser.py:84-92: the ``ae.*`` entries of a FlowPred checkpoint's ``state_dict``.  Returns False (weights stay
    as initialised) when no file is given -- there is no wandb download here (no network in this deployment)."""
    if not path:
        return False
    blob = torch.load(path, map_location="cpu")
    sd = blob["state_dict"] if "state_dict" in blob else blob
    sd = {k[len("ae."):]: v for k, v in sd.items() if k.startswith("ae.")} or sd
    ae.load_state_dict(sd)
    return True
