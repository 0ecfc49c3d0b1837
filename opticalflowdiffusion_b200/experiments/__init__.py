"""Experiment registry, same shape as the reference's experiments/__init__.py:11-30."""
from .exp_matrix_flow import MatrixFlowExperiment

exp_registry = dict(matrix_flow=MatrixFlowExperiment)


def build_experiment(cfg, logger=None, ckpt_path=None):
    return exp_registry[cfg.experiment.name](cfg, logger, ckpt_path)
