"""CPU: the file / wire formats either side of the hot path against goldens produced by the
reference's own ``load_flow`` and ``InputPadder`` (oracle/make_goldens_io.py), plus checkpoint alias handling."""
import os

import numpy as np
import pytest
import torch

from opticalflowdiffusion_b200 import io_formats as IO


def test_flo_reader_matches_reference(golden, tmp_path):
    g = golden("io_formats")
    path = tmp_path / "a.flo"
    path.write_bytes(g["flo_bytes"].tobytes())
    arr = IO.read_flo(str(path))
    assert arr.dtype == np.float32 and np.array_equal(arr, g["flo_array"])          # bit-exact
    out = tmp_path / "b.flo"
    IO.write_flo(str(out), arr)
    assert out.read_bytes() == g["flo_bytes"].tobytes()                              # writer is the exact inverse
    IO.write_flo(str(out), torch.from_numpy(arr).permute(2, 0, 1))                   # (2,h,w) tensors too
    assert out.read_bytes() == g["flo_bytes"].tobytes()
    bad = bytearray(g["flo_bytes"].tobytes())
    bad[0] ^= 0xFF
    path.write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        IO.read_flo(str(path))
    path.write_bytes(g["flo_bytes"].tobytes()[:-8])
    with pytest.raises(ValueError):
        IO.read_flo(str(path))


def test_input_padder_matches_reference(golden):
    g = golden("io_formats")
    for mode_id, h, w, *pad in g["pads"].tolist():
        p = IO.InputPadder((1, 3, h, w), mode="sintel" if mode_id == 0 else "kitti")
        assert p._pad == pad, (mode_id, h, w)
        assert (h + pad[2] + pad[3]) % 8 == 0 and (w + pad[0] + pad[1]) % 8 == 0
    x = torch.arange(2 * 3 * 17 * 23, dtype=torch.float32).reshape(2, 3, 17, 23)
    for mode in ("sintel", "kitti"):
        p = IO.InputPadder(x.shape, mode=mode)
        (xp,) = p.pad(x)
        assert np.array_equal(xp.numpy(), g[f"padded_{mode}"])
        assert torch.equal(p.unpad(xp), x)


def _algo(seed):
    from opticalflowdiffusion_b200 import FlowDiffuser
    from opticalflowdiffusion_b200.config import compose
    torch.manual_seed(seed)
    return FlowDiffuser(compose(["algorithm.target=joint"]).algorithm)


def test_checkpoint_alias_families(tmp_path):
    src, dst = _algo(1), _algo(2)
    path = str(tmp_path / "m.ckpt")
    IO.save_checkpoint(src, path, global_step=7)
    ck = torch.load(path, weights_only=False)
    assert ck["global_step"] == 7 and set(ck["state_dict"]) == set(src.state_dict())
    # full Lightning checkpoint (all aliases)
    rep = IO.load_checkpoint(dst, path)
    assert rep == {"missing": [], "unexpected": []}
    for (k, a), b in zip(src.state_dict().items(), dst.state_dict().values()):
        assert torch.equal(a, b), k
    # a checkpoint that kept only ONE alias family of the UNet still loads
    for family in ("unet.", "model.model.model."):
        dst = _algo(3)
        sd = {k: v for k, v in ck["state_dict"].items()
              if k.startswith(family) or not any(k.startswith(p) for p in IO._UNET_PREFIXES)}
        assert any(k.startswith(family) for k in sd)
        IO.load_checkpoint(dst, {"state_dict": sd})
        assert torch.equal(dst.unet.final_res_block.block1.proj.weight, src.unet.final_res_block.block1.proj.weight)
        assert torch.equal(dst.model.betas, src.model.betas)
    # disagreeing aliases and unknown keys are errors
    sd = dict(ck["state_dict"])
    sd["unet.init_conv.bias"] = sd["unet.init_conv.bias"] + 1
    with pytest.raises(ValueError):
        IO.load_checkpoint(_algo(4), {"state_dict": sd})
    sd = dict(ck["state_dict"])
    sd["something.else"] = torch.zeros(1)
    with pytest.raises(KeyError):
        IO.load_checkpoint(_algo(4), {"state_dict": sd})
    assert IO.load_checkpoint(_algo(4), {"state_dict": sd}, strict=False)["unexpected"] == ["something.else"]


def test_sintel_dataset_from_a_tree(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from opticalflowdiffusion_b200.config import Config
    root = tmp_path / "MPI_Sintel"
    rng = np.random.default_rng(0)
    for scene in ("alley_1", "market_2"):
        (root / "training" / "clean" / scene).mkdir(parents=True)
        (root / "training" / "flow" / scene).mkdir(parents=True)
        for i in range(1, 4):
            img = rng.integers(0, 255, (12, 20, 3), dtype=np.uint8)
            cv2.imwrite(str(root / "training" / "clean" / scene / f"frame_{i:04d}.png"), img)
            if i < 3:
                IO.write_flo(str(root / "training" / "flow" / scene / f"frame_{i:04d}.flo"),
                             rng.standard_normal((12, 20, 2)).astype(np.float32))
    cfg = Config.wrap({"root": str(root), "render": "clean", "image_size": None, "val_fraction": 0.5})
    tr, va = IO.SintelFlowDataset(cfg, "training"), IO.SintelFlowDataset(cfg, "validation")
    assert len(tr) + len(va) == 4 and len(tr) in (0, 2, 4)
    ds = tr if len(tr) else va
    img, tgt, flow = ds[0]
    assert img.shape == (3, 12, 20) and tgt.shape == (3, 12, 20) and flow.shape == (2, 12, 20)
    assert 0.0 <= float(img.min()) and float(img.max()) <= 1.0
    assert np.array_equal(flow.permute(1, 2, 0).numpy(), IO.read_flo(ds.items[0][2]))
    cfg2 = Config.wrap({"root": str(root), "render": "clean", "image_size": "10,6", "val_fraction": 0.5})
    ds2 = IO.SintelFlowDataset(cfg2, "training" if len(tr) else "validation")
    img, tgt, flow = ds2[0]
    assert img.shape == (3, 6, 10) and flow.shape == (2, 6, 10)
