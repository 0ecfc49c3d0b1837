"""CLI with the reference's calling convention (main.py:24-89):
    python -m opticalflowdiffusion_b200.main experiment=matrix_flow algorithm=flow_diffuser \\
        algorithm.target=flow algorithm.sampling_timesteps=50 +experiment.tasks=[validation]
"""
from __future__ import annotations

import json
import sys

import torch

from .config import compose
from .experiments import build_experiment


def run(cfg):
    torch.set_float32_matmul_precision("high")     # main.py:79-80
    exp = build_experiment(cfg, None, cfg.get("ckpt_path"))
    results = {}
    for task in cfg.experiment.tasks:
        results[task] = exp.exec_task(task)
    return results


def main(argv=None):
    cfg = compose(sys.argv[1:] if argv is None else argv)
    print(json.dumps(run(cfg), indent=1, default=str))


if __name__ == "__main__":
    main()
