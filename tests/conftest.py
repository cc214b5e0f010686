import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


if os.environ.get("HB200_EMU_TEST"):
    # child process of tests/test_emu_kernels.py: the gpu-marked parity tests against the host
    # emulation of the kernels (tests/emu_env.py); never set on a GPU box
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_env
    emu_env.activate()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a box without a GPU
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
