#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY -- builds the one native component of the reference that lies on the hot path,
its double-backward CUDA extension (third_party/cuda_gridsample_grad2/gridsample_cuda.{cpp,cu}), UNMODIFIED, from the
sources where they lie under /root/reference, into oracle/_ref/gridsample_grad2.so (git-ignored; it travels to the
GPU box with the snapshot).  nvcc cross-compiles for sm_100a without a GPU; the build needs the ATen headers and
takes ~5 minutes.  Nothing is copied from the reference tree: only the built .so lands here.

Used by tests/test_gpu_gridsample.py (value parity of miso_grid_sample3d_bwd_bwd against the reference's own kernel)
and benchmarks/interp_sweep.py (head-to-head timing).  Run in the build container:  python oracle/build_ref.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("MISO_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "third_party", "cuda_gridsample_grad2")
SO = os.path.join(OUT, "gridsample_grad2.so")


def build(verbose=True):
    if os.path.exists(SO):
        return SO
    if not os.path.isdir(SRC):
        return None   # GPU box: the prebuilt .so is all there is
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="gridsample_grad2", sources=[os.path.join(SRC, "gridsample_cuda.cpp"), os.path.join(SRC, "gridsample_cuda.cu")],
         build_directory=OUT, extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], verbose=verbose,
         is_python_module=False)
    return SO if os.path.exists(SO) else None


def load_module():
    """Import the prebuilt extension (pybind11 module exposing grad2_2d / grad2_3d); None when it was never built."""
    if not os.path.exists(SO):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location("gridsample_grad2", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build() or "reference sources not found; nothing built")
