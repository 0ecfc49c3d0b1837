"""``experiment=matrix_flow`` runner for ``algorithm=flow_diffuser``: the slice of
experiments/exp_base.py:47-59,120-238 + exp_99.py:18-45 that drives the hot path.

The reference builds a ``pl.Trainer`` (DDP when more than one GPU is visible, exp_base.py:198).
Lightning is not part of this path's contract, so the loop here is explicit: one process per GPU
(torchrun), batch-sharded data, ``validation`` / ``sample`` tasks run the CUDA sampling path;
``train`` needs the backward kernels and says so instead of silently doing something else."""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from ..datasets import SyntheticSintelDataset
from ..flow_diffuser import FlowDiffuser
from ..parallel import reduce_metrics


class MatrixFlowExperiment:
    compatible_algorithms = dict(flow_diffuser=FlowDiffuser)
    compatible_datasets = dict(synthetic_sintel=SyntheticSintelDataset)

    def __init__(self, cfg, logger=None, ckpt_path: Optional[str] = None):
        self.cfg, self.logger, self.ckpt_path = cfg, logger, ckpt_path
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        self.algo = self._build_model()

    def _build_model(self):
        name = self.cfg.algorithm.name
        if name not in self.compatible_algorithms:
            raise ValueError(f"algorithm '{name}' is outside the flow_diffuser hot path "
                             f"(have: {sorted(self.compatible_algorithms)})")
        algo = self.compatible_algorithms[name](self.cfg.algorithm)
        if self.ckpt_path:
            sd = torch.load(self.ckpt_path, map_location="cpu")
            algo.load_state_dict(sd.get("state_dict", sd))
        algo.logger = self.logger
        return algo

    def _build_dataset(self, split: str):
        if split not in ("training", "validation", "test"):
            raise NotImplementedError(f"split '{split}' is not implemented")
        return self.compatible_datasets[self.cfg.dataset.name](self.cfg.dataset, split=split)

    def _loader(self, split: str, section):
        ds = self._build_dataset(split)
        sampler = None
        if self.world > 1:      # batch-sharded data parallelism: each rank sees a disjoint slice, no collective
            sampler = torch.utils.data.DistributedSampler(ds, num_replicas=self.world, rank=self.rank,
                                                          shuffle=bool(section.data.shuffle))
        return torch.utils.data.DataLoader(ds, batch_size=int(section.data.batch_size), sampler=sampler,
                                           shuffle=bool(section.data.shuffle) and sampler is None, num_workers=0)

    def exec_task(self, task: str):
        if task == "train":
            return self.train()
        if task in ("validation", "validate", "test"):
            return self.validate()
        raise ValueError(f"Specified task '{task}' not implemented for class {self.__class__.__name__}.")

    def train(self):
        raise NotImplementedError(
            "training needs the dgrad/wgrad/normalisation backward kernels, which are the next build step; "
            "there is deliberately no autograd/PyTorch fallback.  training_step() computes the forward loss.")

    @torch.no_grad()
    def validate(self, max_batches: Optional[int] = None):
        dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(dev)
        self.algo.to(dev)
        limit = max_batches or int(self.cfg.experiment.validation.limit_batch)
        out = {}
        for i, batch in enumerate(self._loader("validation", self.cfg.experiment.validation)):
            if i >= limit:
                break
            self.algo.validation_step(tuple(t.to(dev) for t in batch), i)
            out = {k: float(v) for k, v in self.algo.logged.items()} if hasattr(self.algo, "logged") else {}
        return reduce_metrics(out, device=dev)     # sync_dist=True semantics for the scalars
