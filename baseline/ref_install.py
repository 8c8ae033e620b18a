"""Unmodified reference model sources for the full-model parity tests and bench legs.

`install()` copies `/root/reference/src/d_fine` (pure Python; the model, criterion, matcher) into
`baseline/_ref/src/d_fine` WITHOUT touching a byte.  `baseline/_ref/` is git-ignored (the reference's
sources never enter this repository's history) but not gpurun-ignored, so the copy travels to the GPU
box like the built `.so`.  `import_reference()` puts `baseline/_ref` on `sys.path` so that
`from src.d_fine.dfine import build_model, build_loss, build_optimizer` resolves exactly as it does
inside the reference's own tree (the reference imports itself as `src.d_fine...`, dfine.py:7).

Nothing under `d-fine-seg_b200/` imports this module: the product patches whatever model object the
caller built (`dfine_b200.patch_model`).  Users are `tests/test_gpu_model.py`, `bench.py`'s
`full_model` legs and `--impl reference`.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"
SUBDIRS = ["src/d_fine"]


def installed() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "d_fine", "dfine.py"))


def install(force: bool = False) -> str:
    """Copy the reference's model package verbatim (only where /root/reference exists: the build
    container).  Returns the install root."""
    if not os.path.isdir(SOURCE):
        if installed():
            return REF_ROOT
        raise RuntimeError(f"{SOURCE} is absent and baseline/_ref holds no prebuilt copy")
    for sub in SUBDIRS:
        src, dst = os.path.join(SOURCE, sub), os.path.join(REF_ROOT, sub)
        if os.path.isdir(dst) and not force:
            cmp = filecmp.dircmp(src, dst, ignore=["__pycache__"])
            if not (cmp.left_only or cmp.diff_files or cmp.funny_files):
                continue
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return REF_ROOT


def import_reference():
    """Make `src.d_fine` importable; returns the module `src.d_fine.dfine`."""
    if not installed():
        install()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib

    return importlib.import_module("src.d_fine.dfine")


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
