"""``FlowLearner`` (reference: algorithms/diffusion_animation/flow_learner.py:61-340, ``algorithm=flow_learner``), flow
representation: a time-free UNet (``Unet(64, channels=6, out_dim=3, time_in=False)``) predicts flow + splat weights from
a frame pair and is trained with the multi-scale soft-splat photometric loss (:133-222) -- SURVEY.md section 8f row N1.

Everything runs on the kernels of the flow_diffuser path: the UNet forward / backward (``UnetFunction``), ``fd_splat_*``
with the reference's ``scale`` / ``offset`` grid, and two fused loss kernels (``fd_soft_charb_*``: normalise + hole fill +
NaN-aware Charbonnier mean of one (level, offset) term; ``fd_edge_smooth_*``).  The filter representation (``radius``,
``ConvToFilter``) is outside the hot path and raises.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import warp as W
from .flow_diffuser import UnetWithWarp, _Base, _cfg_get
from .unet import Unet

Tensor = torch.Tensor

LEVELS = (1, 2, 4, 5, 7, 8, 10, 11, 14, 16)      # flow_learner.py:167


class FlowLearner(_Base):
    def __init__(self, cfg, levels: Sequence[int] = LEVELS):
        super().__init__()
        self.cfg = cfg
        if _cfg_get(cfg, "radius") is not None:
            raise NotImplementedError("the filter representation (radius / ConvToFilter) is outside the flow hot path")
        self.radius = None
        self.flow_max = cfg.flow_max
        self.rep = "flow"
        self.levels = tuple(levels)
        self.fused_levels = True            # False: one launch set per (level, a, b) like the reference's loop
        self._augmentor = None
        # 3 outputs: optical flow + splat weight map (:88-93)
        self.unet = UnetWithWarp(cfg, Unet(64, channels=6, out_dim=3, time_in=False), False, nan_safe=False)
        self.model = self.unet

    @property
    def augmentor(self):
        if self._augmentor is None:
            from .augment import GpuAugmentor
            self._augmentor = GpuAugmentor()
        return self._augmentor

    def configure_optimizers(self):
        from .optim import FusedAdam

        def invalidate():
            self.unet.model.weights_epoch += 1

        self.optimizers = FusedAdam(self.model.parameters(), lr=self.cfg.lr, weight_decay=self.cfg.weight_decay,
                                    on_step=invalidate)
        return self.optimizers

    def preprocess(self, batch, aug: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
        """:109-125: (tgt, cat(img, tgt), flow) with frames in [-1, 1] and the flow normalised by flow_max."""
        if aug:
            batch = self.augmentor(batch)
        img, tgt, flow = batch
        flow = torch.clamp(flow / self.flow_max, -1.0, 1.0)
        img = 2 * img - 1.0
        tgt = 2 * tgt - 1.0
        return tgt, torch.cat((img, tgt), dim=1), flow

    # ------------------------------------------------------------------ loss
    def _predict(self, cond: Tensor):
        out = self.model(cond, additional_out=True)
        fw = out[:, -3:]
        return fw[:, :2] * self.flow_max, fw[:, 2:]

    def loss(self, tgt: Tensor, cond: Tensor, flow_: Tensor, override_flow=None) -> Tensor:
        """:133-222.  For every level and every (a, b) offset of that level: soft-splat the input frame along the
        predicted flow and the target along zero flow into the (H/level, W/level) grid, NaN the holes, Charbonnier;
        mean over offsets, then over levels, plus 0.01 x edge-aware smoothness."""
        if override_flow is None:
            flow_pred, warp_weights = self._predict(cond)
        else:
            flow_pred = override_flow * self.flow_max
            warp_weights = torch.ones_like(flow_pred[:, :1])
        return self.objective(cond[:, :3].contiguous(), tgt, flow_pred, warp_weights)

    def objective(self, input_img: Tensor, tgt: Tensor, flow_pred: Tensor, warp_weights: Tensor) -> Tensor:
        """The loop of :160-205 given the prediction (flow in pixels, splat weight map)."""
        e = warp_weights.exp()
        ten_in = torch.cat([input_img * e, e], 1)                       # softsplat 'soft' input (softsplat_new.py:306-307)
        with torch.no_grad():
            et = torch.full_like(tgt[:, :1], 1.0).exp()
            tgt_in = torch.cat([tgt * et, et], 1).contiguous()
            zero_flow = torch.zeros_like(flow_pred)
        photo = []
        for level in self.levels:
            if self.fused_levels:
                photo.append(W.soft_level_loss(ten_in, flow_pred, tgt_in, level))       # all level^2 offsets in one go
                continue
            terms = []
            for a in range(level):
                for b in range(level):
                    S = W.softsplat_func.apply(ten_in, flow_pred, level, a, b)
                    with torch.no_grad():
                        T = W.softsplat_func.apply(tgt_in, zero_flow, level, a, b)
                    terms.append(W.soft_splat_charbonnier(S, T))
            photo.append(torch.stack(terms).mean())
        loss = torch.stack(photo).mean()
        return loss + 0.01 * W.edgeaware_smoothness1(input_img, flow_pred)

    @torch.no_grad()
    def sample(self, cond: Tensor, flo: Tensor, log_additional: bool = False):
        """:224-240: (samples, flow in pixels, splat weights)."""
        flow, warp_weights = self._predict(cond)
        sw = W.softsplat(cond[:, :3].contiguous(), flow.contiguous(), warp_weights.contiguous(), "soft", scale=1, offset=(0, 0))
        samples = W.fill_holes_nan(sw[:, :-1], sw[:, -1:])
        return samples, flow, warp_weights

    # ------------------------------------------------------------------ steps
    @staticmethod
    def _stats(prefix: str, name: str, x: Tensor):
        from .flow_diffuser import logged_stats
        return logged_stats(prefix, name, x)

    def training_step(self, batch, batch_idx):
        tgt, cond, flow = self.preprocess(batch, aug=bool(_cfg_get(self.cfg, "train_aug", True)))
        loss = self.loss(tgt, cond, flow)
        self.log_dict({"train/loss": loss, **self._stats("train", "cond", cond), **self._stats("train", "flow", flow)})
        self.log("loss", loss, prog_bar=True)
        return loss

    @torch.no_grad()
    def validation_step(self, batch, batch_idx):
        """:260-340 (scalar metrics)."""
        img, tgt, flow = batch
        tgt_, cond, flow_ = self.preprocess(batch, aug=False)
        loss = self.loss(tgt_, cond, flow_)
        ideal = self.loss(tgt_, cond, flow_, override_flow=flow_)
        samples, p_flows, _ = self.sample(cond, flow_)
        samples = torch.where(torch.isnan(samples), torch.zeros_like(samples), samples)
        metrics = {"val/loss": loss, "val/ideal_loss": ideal, "val/mse": torch.nn.functional.mse_loss(samples, tgt),
                   "val/flow_mse": torch.nn.functional.mse_loss(flow_, p_flows / self.flow_max),
                   **self._stats("val", "cond", cond), **self._stats("val", "flow", flow),
                   **self._stats("val", "samples", samples), **self._stats("val", "p_flow", p_flows)}
        self.log_dict(metrics, sync_dist=True)
        return None
