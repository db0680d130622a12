import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "umi-collapse-rs_b200"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) and libumigpu.so (nvcc) if they are not there yet."""
    if not os.path.exists(os.path.join(REPO, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")])
    if not os.path.exists(os.path.join(REPO, "umi-collapse-rs_b200", "csrc", "libumigpu.so")):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "umi-collapse-rs_b200", "csrc")])
