"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the data formats either side of the hot path,
produced by the UNMODIFIED reference code where it lies under /root/reference:

* ``SintelDataset.load_flow`` (datasets/animation/sintel.py:59-65) -- the method's source is lifted out of the file with
  ``ast`` (the module itself imports matplotlib / omegaconf, absent here; the method only needs numpy) and run on a
  ``.flo`` file assembled byte by byte with ``struct``;
* ``InputPadder`` (algorithms/diffusion_animation/future/raft_utils.py:7-25) imported directly.

Run in the build container:  python oracle/make_goldens_io.py   ->  tests/golden/io_formats.npz
"""
import ast
import importlib.util
import os
import struct
import sys
import tempfile
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FLOWDIFF_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def reference_load_flow():
    src = open(os.path.join(REF, "datasets", "animation", "sintel.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "load_flow":
            code = textwrap.dedent(ast.get_source_segment(src, node))
            ns = {"np": np}
            exec(code, ns)
            return ns["load_flow"]
    raise RuntimeError("load_flow not found in the reference")


def reference_input_padder():
    path = os.path.join(REF, "algorithms", "diffusion_animation", "future", "raft_utils.py")
    spec = importlib.util.spec_from_file_location("ref_raft_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.InputPadder


def main():
    d = {}
    # ---- .flo
    rng = np.random.default_rng(5)
    h, w = 7, 11
    flow = rng.standard_normal((h, w, 2)).astype(np.float32) * 3
    raw = struct.pack("<f", 202021.25) + struct.pack("<ii", w, h) + flow.tobytes()
    with tempfile.NamedTemporaryFile(suffix=".flo", delete=False) as f:
        f.write(raw)
        path = f.name
    load_flow = reference_load_flow()
    got = load_flow(None, path)
    os.unlink(path)
    d["flo_bytes"] = np.frombuffer(raw, np.uint8)
    d["flo_array"] = np.asarray(got, np.float32)
    # ---- InputPadder
    Padder = reference_input_padder()
    dims = [(436, 1024), (370, 1226), (64, 128), (375, 1242), (17, 23), (8, 8), (9, 15)]
    pads = []
    for mode in ("sintel", "kitti"):
        for hh, ww in dims:
            p = Padder((1, 3, hh, ww), mode=mode)
            pads.append([0 if mode == "sintel" else 1, hh, ww] + list(p._pad))
    d["pads"] = np.array(pads, np.int64)
    x = torch.arange(2 * 3 * 17 * 23, dtype=torch.float32).reshape(2, 3, 17, 23)
    for mode in ("sintel", "kitti"):
        p = Padder(x.shape, mode=mode)
        (xp,) = p.pad(x)
        d[f"padded_{mode}"] = xp.numpy()
        assert torch.equal(p.unpad(xp), x)
    np.savez_compressed(os.path.join(GOLD, "io_formats.npz"), **d)
    print("io_formats.npz", {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
