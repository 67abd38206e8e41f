#!/usr/bin/env python
"""Build the reference's OWN CUDA kernel for the quantize entry point as a timing comparator.

TEST / BENCH INFRASTRUCTURE ONLY (nothing under mcaq_yolo_b200/ may load it).  The source is compiled
where it lies -- /root/reference/mcaq_yolo/ops/src/mcaq_kernel.cu (one file, no torch, no other
dependency) -- with nvcc for sm_100, output only into oracle/_ref/ (git-ignored, travels to the GPU box
with the snapshot).  No reference source is copied into this repository.

It is "the kernel to beat" for the Level-0 entry point, NOT an oracle: it differs from the
reference's PyTorch path (roundf instead of round-half-even, h / tile_h instead of the nearest rule), which is
what parity is defined against.  tools/ref_kernel_bench.py times it beside libmcaq_b200.so.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libmcaq_ref_kernel.so")
# C++ linkage in the reference (mcaq_kernel.cu:102): the Itanium-mangled name of
# launch_spatial_quantization(const float*, const float*, const float*, const float*, const float*, float*,
#                             int x8, CUstream_st*)
SYMBOL = "_Z27launch_spatial_quantizationPKfS0_S0_S0_S0_PfiiiiiiiiP11CUstream_st"


def source():
    for root in (os.environ.get("MCAQ_REF"), "/root/reference"):
        if root:
            p = os.path.join(root, "mcaq_yolo", "ops", "src", "mcaq_kernel.cu")
            if os.path.exists(p):
                return p
    return None


def build(force=False):
    """Returns the path of the built library, or None when the reference tree is absent (GPU box: the
    prebuilt file from the snapshot is used)."""
    src = source()
    if src is None:
        return OUT if os.path.exists(OUT) else None
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-gencode", "arch=compute_100,code=sm_100", "-O3", "-shared", "-Xcompiler", "-fPIC", "-o", OUT, src]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
