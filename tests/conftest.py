import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    """CUDA devices visible to the runtime the product library links against (0 without a driver)."""
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import glob
            cands = sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))
            rt = ctypes.CDLL(cands[-1]) if cands else None
        except OSError:
            rt = None
    if rt is None:
        return 0
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def pytest_collection_modifyitems(config, items):
    """Plain `pytest` on a machine without a CUDA device: the @pytest.mark.gpu items are skipped, not failed (the
    product has no CPU fallback).  On a GPU machine a missing libbmpc.so is NOT a reason to skip: the tests fail."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the CUDA path has no CPU fallback")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
