"""TEST INFRASTRUCTURE ONLY -- compile the reference's forward-splat kernels for the HOST.

The reference's only native code is three CUDA kernels kept as Python strings inside
``algorithms/diffusion_animation/softsplat_new.py`` (``softsplat_out`` :352-423,
``softsplat_ingrad`` :489-565, ``softsplat_flowgrad`` :600-700) and JIT-compiled by CuPy,
which is absent here; the Python wrapper refuses CPU tensors (:444-445).  To obtain outputs
of the *reference's own arithmetic* without a GPU this script

  1. parses ``softsplat_new.py`` where it lies under ``/root/reference`` (``ast``; nothing is
     copied into the repo) and pulls out the three kernel-source string literals,
  2. specialises them with the reference's own templating function ``cuda_kernel()``
     (:31-249 -- pure Python; shapes and strides are baked in as literals per call),
  3. prepends a 12-line host shim (``blockIdx``/``blockDim`` = one sequential "thread",
     ``atomicAdd`` = ``+=``, ``return`` = ``continue`` inside the grid-stride loop) and compiles
     with ``g++ -ffp-contract=off`` into ``oracle/_ref/`` (git-ignored, not gpurun-ignored).

The resulting functions execute the reference kernels' statements verbatim, element by element.
They are used by ``oracle/make_goldens.py`` to produce ``tests/golden/splat_*.npz`` and by
``tests/test_oracle_vs_reference.py`` (skipped when ``/root/reference`` is absent).
"""
from __future__ import annotations

import ast
import ctypes
import hashlib
import os
import subprocess

import torch

from . import ref_stubs

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(_HERE, "_ref")

_SHIM = r"""
#include <cmath>
#include <cstdlib>
#include <cassert>
using std::isfinite; using std::floor; using std::abs;
struct fd_dim3 { int x, y, z; };
static const fd_dim3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
template <class T> static inline void atomicAdd(T* p, T v) { *p += v; }
#define __global__
#define __launch_bounds__(x)
#define return continue
"""

_kernel_sources = None


def _extract_kernel_sources():
    """{'softsplat_out': src, ...} from the reference file's AST (calls to cuda_kernel(name, src, vars))."""
    global _kernel_sources
    if _kernel_sources is not None:
        return _kernel_sources
    path = os.path.join(ref_stubs.REFERENCE_ROOT, "algorithms", "diffusion_animation", "softsplat_new.py")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = ast.parse(open(path).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "id", None) == "cuda_kernel" and len(node.args) >= 2:
            a0, a1 = node.args[0], node.args[1]
            if isinstance(a0, ast.Constant) and isinstance(a1, ast.Constant):
                found[a0.value] = a1.value
    assert set(found) == {"softsplat_out", "softsplat_ingrad", "softsplat_flowgrad"}, sorted(found)
    _kernel_sources = found
    return found


def _compile(name: str, variables: dict):
    ns = ref_stubs.import_reference()
    ss = ns.softsplat_new
    ss.objCudacache.setdefault("device", "host-shim")
    key = ss.cuda_kernel(name, _extract_kernel_sources()[name], variables)
    src = ss.objCudacache[key]["strKernel"]
    digest = hashlib.sha1(src.encode()).hexdigest()[:16]
    os.makedirs(REF_OUT, exist_ok=True)
    so = os.path.join(REF_OUT, f"{name}_{digest}.so")
    if not os.path.exists(so):
        cpp = os.path.join(REF_OUT, f"{name}_{digest}.cpp")
        with open(cpp, "w") as f:
            f.write(_SHIM + src)
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-w", "-shared", "-fPIC",
                               "-o", so, cpp])
    return getattr(ctypes.CDLL(so), name)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def ref_splat_out(ten_in: torch.Tensor, flow: torch.Tensor, scale: int, off_x: int, off_y: int) -> torch.Tensor:
    """softsplat_func.forward launch logic (softsplat_new.py:343-345, 428-438) on the host build."""
    ten_in = ten_in.contiguous().float()
    flow = flow.contiguous().float()
    out = ten_in.new_zeros([ten_in.shape[0], ten_in.shape[1], ten_in.shape[2] // scale, ten_in.shape[3] // scale])
    fn = _compile("softsplat_out", {"tenIn": ten_in, "tenFlow": flow, "tenOut": out})
    fn(ctypes.c_int(ten_in.nelement()), _ptr(ten_in), _ptr(flow), _ptr(out),
       ctypes.c_int(scale), ctypes.c_int(off_x), ctypes.c_int(off_y))
    return out


def ref_splat_backward(ten_in, flow, outgrad, scale: int, off_x: int, off_y: int):
    """softsplat_func.backward launch logic (softsplat_new.py:460-730): (ingrad, flowgrad)."""
    ten_in = ten_in.contiguous().float()
    flow = flow.contiguous().float()
    outgrad = outgrad.contiguous().float()
    ingrad = torch.zeros_like(ten_in)
    flowgrad = torch.zeros_like(flow)
    variables = {"tenIn": ten_in, "tenFlow": flow, "tenOutgrad": outgrad, "tenIngrad": ingrad,
                 "tenFlowgrad": flowgrad}
    f1 = _compile("softsplat_ingrad", variables)
    f1(ctypes.c_int(ingrad.nelement()), _ptr(ten_in), _ptr(flow), _ptr(outgrad), _ptr(ingrad), _ptr(None),
       ctypes.c_int(scale), ctypes.c_int(off_x), ctypes.c_int(off_y))
    f2 = _compile("softsplat_flowgrad", variables)
    f2(ctypes.c_int(flowgrad.nelement()), _ptr(ten_in), _ptr(flow), _ptr(outgrad), _ptr(None), _ptr(flowgrad),
       ctypes.c_int(scale), ctypes.c_int(off_x), ctypes.c_int(off_y))
    return ingrad, flowgrad


def build():
    """Pre-build the shapes the goldens use (called from __graft_entry__.build() when the
    reference tree is present)."""
    if not ref_stubs.reference_available():
        return False
    t = torch.zeros(1, 4, 8, 8)
    f = torch.zeros(1, 2, 8, 8)
    ref_splat_out(t, f, 1, 0, 0)
    ref_splat_backward(t, f, torch.zeros(1, 4, 8, 8), 1, 0, 0)
    return True


if __name__ == "__main__":
    print("built" if build() else "reference tree absent; nothing built")
