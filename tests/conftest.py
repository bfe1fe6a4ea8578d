import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
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


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library and the CPU test harness if they are missing."""
    lib = os.path.join(ROOT, "ectrans_b200", "lib", "libectrans_b200.so")
    emu = os.path.join(ROOT, "tests", "hostemu", "_build", "libemu.so")
    if not (os.path.exists(lib) and os.path.exists(emu)):
        subprocess.check_call(["make", "-C", ROOT, "-j8", "all", "emu"])
    return lib, emu


@pytest.fixture(scope="session")
def emu(built):
    import ctypes
    return ctypes.CDLL(built[1])


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    return {
        "nloen": np.load(os.path.join(g, "lon_number_by_lat.npy")),
        "nmen": np.load(os.path.join(g, "zonal_wavenumbers.npy")),
        "sp": np.load(os.path.join(g, "tl149-c24-s1t@sp.npy")),
        "gp_latlon": np.load(os.path.join(g, "tl149-c24-s1t@sp2gp.npy")),
    }
