"""B200-native (sm_100a) implementation of the ``flow_diffuser`` hot path of
davidfang00/opticalflowdiffusion: noise-prediction UNet, DDPM/DDIM scheduler, backward warp /
forward splat and the photometric / EPE / NaN-MSE losses, behind the reference's algorithm class.

The arithmetic lives in ``lib/libflowdiff.so`` (C ABI: ``include/flowdiff.h``); Python/PyTorch is
plumbing (device memory, streams, ``torch.distributed``).  Build: ``python -m opticalflowdiffusion_b200.build``.
"""
__version__ = "0.1.0"

__all__ = ["FlowDiffuser", "FlowLearner", "Unet", "ConditionalDiffusion", "compose"]


def __getattr__(name):      # lazy: importing the package must not need the built library
    if name == "FlowDiffuser":
        from .flow_diffuser import FlowDiffuser
        return FlowDiffuser
    if name == "FlowLearner":
        from .flow_learner import FlowLearner
        return FlowLearner
    if name == "Unet":
        from .unet import Unet
        return Unet
    if name == "ConditionalDiffusion":
        from .diffusion import ConditionalDiffusion
        return ConditionalDiffusion
    if name == "compose":
        from .config import compose
        return compose
    raise AttributeError(name)
