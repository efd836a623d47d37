import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _ensure_built():
    """The CUDA library is git-ignored: build it (nvcc cross-compiles without a GPU) when a fresh checkout runs the tests
    before __graft_entry__.build()."""
    import shutil
    import subprocess

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    so = os.path.join(root, "seal_b200", "libseal_b200.so")
    if not os.path.exists(so) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        subprocess.run(["make", "-C", os.path.join(root, "seal_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def pytest_configure(config):
    _ensure_built()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
