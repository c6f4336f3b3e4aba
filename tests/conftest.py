import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "spacecraft-pose-estimation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if os.environ.get("SPE_TEST_LIB"):  # A/B runs of the GPU tests against another build (tools/ab_builds.py); the product reads no environment
        from spe_b200 import _lib

        _lib.LIB_PATH = os.path.abspath(os.environ["SPE_TEST_LIB"])


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def decode_golden():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "decode_golden.npz"))


@pytest.fixture(scope="session")
def pnp_golden():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "pnp_golden.npz"))
