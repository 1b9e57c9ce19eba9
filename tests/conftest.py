import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # On a GPU-less machine the gpu tests are skipped; on a GPU box they run and FAIL (never skip)
    # if libarap_b200.so is missing, because the engine has no fallback.
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (gpu tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def meshes():
    z = np.load(os.path.join(GOLDEN, "meshes.npz"))
    return {n: (z[n + "_V"], z[n + "_F"]) for n in ("bar", "sphere", "plane")}


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "arap_golden.npz"))


@pytest.fixture(scope="session")
def trajectory_golden():
    return np.load(os.path.join(GOLDEN, "trajectory_golden.npz"))


def bbox_diag(P):
    return float(np.linalg.norm(P.max(0) - P.min(0)))
