import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # Build what is missing (the driver normally runs __graft_entry__.build() first; this keeps a bare
    # `pytest` usable).  Built .so files travel to the GPU box, where nothing is rebuilt.
    from tesseract_decoder_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tesseract_decoder_b200", "csrc"), "lib"], check=True)
    from oracle import oracle
    if not oracle.available("port"):
        oracle.build("port")
    if not oracle.available("reference") and os.path.isdir("/root/reference/src"):
        oracle.build("reference")


def _has_gpu() -> bool:
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        cudart = None
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return cudart is not None and os.path.exists("/dev/nvidia0")


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_kinds():
    from oracle import oracle
    return [k for k in ("reference", "port") if oracle.available(k)]
