"""pytest configuration: registers the `gpu` marker and puts the package directory on sys.path.

`-m "not gpu"` runs on the GPU-less build box (oracle vs golden vectors, host logic, C-ABI symbol check);
`-m gpu` are the parity tests proper and call the CUDA kernels through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pvsr_lib():
    sys.path.insert(0, os.path.join(PKG, "csrc"))
    import build as pvsr_build
    pvsr_build.build()
    from pvsr import lib
    return lib.load()
