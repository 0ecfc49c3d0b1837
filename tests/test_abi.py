"""The C-ABI library loads and exports every symbol include/flowdiff.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

from opticalflowdiffusion_b200 import _lib
from opticalflowdiffusion_b200 import build as fd_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "flowdiff.h")).read()
    return sorted(set(re.findall(r"FD_API\s+[\w\s\*]+?\b(fd_\w+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    assert len(syms) >= 30
    for must in ("fd_conv_igemm", "fd_backwarp_fwd", "fd_backwarp_photo_epe_fwd", "fd_splat_fwd", "fd_splat_flowgrad",
                 "fd_ddim_step", "fd_ddpm_step", "fd_q_sample", "fd_nan_mse_fwd", "fd_gn_silu", "fd_linattn", "fd_attention"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    path = fd_build.build()
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.fd_arch() == b"sm_100a" and lib.fd_version() >= 100


def test_no_cpu_fallback():
    """Ops refuse CPU tensors instead of silently computing elsewhere."""
    import pytest
    import torch
    from opticalflowdiffusion_b200 import warp
    with pytest.raises(_lib.FlowDiffError):
        warp.warp_backward_flow(None, torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4))
    from opticalflowdiffusion_b200.unet import Unet
    with pytest.raises(_lib.FlowDiffError):
        Unet(64, 5, 2)(torch.zeros(1, 2, 8, 8), torch.zeros(1, 3, 8, 8), torch.zeros(1, dtype=torch.long))
