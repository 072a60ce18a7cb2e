"""Build the C-ABI shared library (CUDA kernels for sm_100a + host flattener) in-tree.

    python scikit-garden_b200/build.py          # or: __graft_entry__.build()

nvcc cross-compiles without a GPU.  The resulting ``libqforest_b200.so`` sits next to
this file (git-ignored, but it travels to the GPU box with the snapshot).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libqforest_b200.so")
SOURCES = ["qforest_abi.cu", "pack_forest.cpp"]
HEADERS = ["apply_kernel.cuh", "quantile_kernel.cuh", "select_kernel.cuh", "select2_kernel.cuh", "select3_kernel.cuh", "select4_kernel.cuh", "compact_lists.cuh", "csr_builder.cuh", "qf_internal.h",
           os.path.join("..", "..", "include", "qforest_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",          # the float32 op order of the reference has no fused multiply-add
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
