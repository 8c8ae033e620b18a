"""In-tree build of libdfine_b200.so with nvcc for sm_100a (no JIT cache, no torch extension).

The shared library is a plain C-ABI object (include/dfine_b200.h): it links only the CUDA
runtime, so it cross-compiles on a GPU-less box and travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
PROJ_DIR = os.path.dirname(PKG_DIR)
REPO_ROOT = os.path.dirname(PROJ_DIR)
CSRC = os.path.join(PROJ_DIR, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")
LIB_DIR = os.path.join(PKG_DIR, "_C")
# DFINE_B200_LIB: load another build of the same C-ABI (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("DFINE_B200_LIB") or os.path.join(LIB_DIR, "libdfine_b200.so")

SOURCES = ["api.cu", "msda_fwd.cu", "msda_bwd.cu", "msda_bwd_value.cu", "fdr.cu", "mask_gemm.cu", "wgrad_gemm.cu", "reduce.cu", "lsap.cu", "mask_loss.cu", "linear_fused.cu", "lqe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build libdfine_b200.so")
    return cand


def sources() -> List[str]:
    return [os.path.join(CSRC, s) for s in SOURCES]


def is_stale() -> bool:
    if os.environ.get("DFINE_B200_LIB"):
        return False          # an explicitly named library is used as it is (and must exist)
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "msda_common.cuh"), os.path.join(CSRC, "tma_util.cuh"), os.path.join(CSRC, "umma_util.cuh"),
            os.path.join(INCLUDE, "dfine_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into dfine_b200/_C/libdfine_b200.so."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that confuse nvcc's host compiler detection
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    objs = []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src).replace(".cu", ".o"))
        cmd = [nvcc, "-ccbin", host, *NVCC_FLAGS,
               *os.environ.get("DFINE_NVCC_EXTRA", "").split(),   # e.g. -DDFINE_BV_PROF (debug builds)
               "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True, env=env)
        objs.append(obj)
    tmp = LIB_PATH + ".tmp"
    subprocess.run([nvcc, "-ccbin", host, "-shared", "-o", tmp, *objs, "-cudart", "static"],
                   check=True, env=env)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
