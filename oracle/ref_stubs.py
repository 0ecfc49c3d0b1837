"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Import shim that lets the *unmodified* reference modules under ``/root/reference``
be imported in the build container.  The reference pulls in
packages that are absent here (pytorch_lightning, omegaconf, cupy, ema_pytorch,
accelerate, pytorch_fid, spatial_correlation_sampler and an un-shipped ``utils``
package: flow_diffuser.py:1-17, denoising_diffusion.py:23-29, softsplat_new.py:4).
None of them take part in the arithmetic of the hot path, so they are replaced by
inert stand-ins placed in ``sys.modules`` *before* the reference is imported.

The reference tree is read-only and does not exist on the GPU box, so this module
is only used by ``oracle/make_goldens.py`` and ``oracle/build_ref.py`` (which run
here) and by tests that skip themselves when ``/root/reference`` is missing.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FLOWDIFF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "algorithms", "diffusion_animation"))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    """Insert stand-in modules.  Idempotent."""
    import torch

    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_fd_stub", False):
        return

    class _LightningModule(torch.nn.Module):
        """nn.Module with the handful of Lightning attributes flow_diffuser.py touches."""

        global_step = 0
        logger = None

        def __init__(self, *a, **k):
            super().__init__()
            self.logged = {}

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        def log_dict(self, d, **kwargs):
            self.logged.update({k: (v.detach() if torch.is_tensor(v) else v) for k, v in d.items()})

        def log(self, k, v, **kwargs):
            self.logged[k] = v

    pl = _module("pytorch_lightning", LightningModule=_LightningModule, _fd_stub=True)
    pl.Trainer = object
    _module("pytorch_lightning.strategies")
    _module("pytorch_lightning.strategies.ddp", DDPStrategy=object)
    _module("pytorch_lightning.loggers")
    _module("pytorch_lightning.loggers.wandb", WandbLogger=object)
    _module("pytorch_lightning.utilities")
    _module("pytorch_lightning.utilities.types", TRAIN_DATALOADERS=object, STEP_OUTPUT=object)
    _module("pytorch_lightning.core")
    _module("pytorch_lightning.core.datamodule", LightningDataModule=object)
    _module("pytorch_lightning.callbacks", LearningRateMonitor=object, ModelCheckpoint=object)

    class _DictConfig(dict):
        __getattr__ = dict.get

    class _OmegaConf:
        @staticmethod
        def to_container(c):
            return dict(c)

        @staticmethod
        def create(d):
            return _DictConfig(d)

    _module("omegaconf", DictConfig=_DictConfig, OmegaConf=_OmegaConf)

    _module("ema_pytorch", EMA=object)
    _module("accelerate", Accelerator=object)
    _module("pytorch_fid")
    _module("pytorch_fid.inception", InceptionV3=object)
    _module("pytorch_fid.fid_score", calculate_frechet_distance=lambda *a, **k: 0.0)
    _module("spatial_correlation_sampler", SpatialCorrelationSampler=object)

    def _memoize(for_each_device=False):
        def deco(fn):
            return fn
        return deco

    # einops probes sys.modules['cupy'] and needs .ndarray to exist
    cupy = _module("cupy", ndarray=type("ndarray", (), {}), int32=int, float32=float, memoize=_memoize)
    cupy.cuda = types.SimpleNamespace(get_cuda_path=lambda: "/usr/local/cuda", compile_with_cache=None)

    utils = _module("utils")
    utils.__path__ = []
    vp = _module("utils.video_prediction")
    vp.__path__ = []
    _module("utils.video_prediction.visualization", log_video=lambda *a, **k: None)
    ip = _module("utils.image_prediction")
    ip.__path__ = []
    _module("utils.image_prediction.logging", log_photos=lambda *a, **k: None)
    _module("utils.wandb_utils",
            download_latest_checkpoint=lambda *a, **k: None,
            rewrite_checkpoint_for_compatibility=lambda p: p)


def import_reference():
    """Return the reference's ``algorithms.diffusion_animation`` sub-modules as a namespace.

    Only the modules on the hot path are imported (the package ``__init__`` would pull
    in PWC-Net / RAFT as well, so it is bypassed by registering empty parent packages).
    """
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    import importlib

    if "algorithms" not in sys.modules or not getattr(sys.modules["algorithms"], "_fd_pkg", False):
        pkg = _module("algorithms", _fd_pkg=True)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "algorithms")]
        sub = _module("algorithms.diffusion_animation")
        sub.__path__ = [os.path.join(REFERENCE_ROOT, "algorithms", "diffusion_animation")]
    ns = types.SimpleNamespace()
    for name in ("softsplat_new", "warp", "losses", "augmentation", "denoising_diffusion", "flow_pred",
                 "flow_diffuser"):
        setattr(ns, name, importlib.import_module(f"algorithms.diffusion_animation.{name}"))
    return ns


def reference_cfg(**over):
    """``cfg.algorithm`` as composed from configurations/algorithm/flow_diffuser.yaml:1-19."""
    install_stubs()
    from omegaconf import DictConfig
    cfg = dict(name="flow_diffuser", image_size=128, latent_dim=16, flow_max=20, latent_max=2, lr=1e-5,
               flow_weight=0.0, weight_decay=1e-6, is_diffusion=True, latent=False, timesteps=1000,
               target="joint", ae="px8q8g0m", noiser="image", zero_init=True)
    cfg.update(over)
    return DictConfig(cfg)
