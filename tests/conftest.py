"""pytest configuration: markers and import paths.

`-m "not gpu"` runs on the build container's CPU (oracle vs golden vectors, host logic,
C-ABI symbol table); `-m gpu` runs the parity tests proper on a B200 through the C-ABI.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "d-fine-seg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
