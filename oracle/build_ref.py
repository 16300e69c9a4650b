"""TEST INFRASTRUCTURE ONLY — builds the reference's OWN CUDA ops as the GPU-side checker for a4 / a5.

The reference's upfirdn2d and fused_bias_act exist only as CUDA kernels behind two pybind modules that it JIT-builds at import
time (modules/psp/stylegan2/op/fused_act.py:8-15, op/upfirdn2d.py:7-14). This recipe compiles those same four source files
WHERE THEY LIE under /root/reference (nothing is copied into the repo) with torch.utils.cpp_extension for sm_100a and leaves only
the two shared objects in oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun like this package's own .so).
Only tests/ load them (tests/test_reference_ops_gpu.py), as the checker, never the product.

    python oracle/build_ref.py        # build container only: /root/reference does not exist on the GPU box
"""
from __future__ import annotations

import os
import shutil
import sys
from pathlib import Path

REF_OPS = Path("/root/reference/modules/psp/stylegan2/op")
OUT = Path(__file__).resolve().parent / "_ref"
MODULES = {"fmi_ref_fused": ["fused_bias_act.cpp", "fused_bias_act_kernel.cu"],
           "fmi_ref_upfirdn2d": ["upfirdn2d.cpp", "upfirdn2d_kernel.cu"]}


def built() -> dict:
    return {name: OUT / f"{name}.so" for name in MODULES if (OUT / f"{name}.so").exists()}


def build(verbose: bool = False) -> dict:
    if not REF_OPS.exists():
        return built()
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    OUT.mkdir(parents=True, exist_ok=True)
    for name, files in MODULES.items():
        so = OUT / f"{name}.so"
        srcs = [REF_OPS / f for f in files]
        if so.exists() and so.stat().st_mtime >= max(s.stat().st_mtime for s in srcs):
            continue
        work = OUT / f"_build_{name}"
        work.mkdir(parents=True, exist_ok=True)
        load(name, sources=[str(s) for s in srcs], build_directory=str(work), verbose=verbose, with_cuda=True,
             extra_cuda_cflags=["-lineinfo"])
        shutil.copy2(work / f"{name}.so", so)
        shutil.rmtree(work, ignore_errors=True)
    return built()


def load_built(name: str):
    """Import oracle/_ref/<name>.so (a pybind11 module: `fused_bias_act(...)` / `upfirdn2d(...)`)."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    so = OUT / f"{name}.so"
    if not so.exists():
        raise FileNotFoundError(f"{so} is missing: run `python oracle/build_ref.py` in the build container")
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
