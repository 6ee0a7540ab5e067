import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


# the library caches its RNB_* knobs at first use; the cases here flip them mid-process (monkeypatch.setenv)
os.environ.setdefault("RNB_ENV_NOCACHE", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_ready():
    try:
        import torch

        from rosnet_b200.engine import lib

        return torch.cuda.is_available() and os.path.exists(lib.LIB_PATH)
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CPU box skips the GPU cases instead of failing 360 of them (the product path has no CPU
    fallback, so they cannot pass there); `-m gpu` on a box without CUDA or without the built library still fails loudly."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _gpu_ready():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built librosnet_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "blocked_tensordot.npz"))


@pytest.fixture(params=["auto", "tensor"])
def contract_route(request, monkeypatch):
    """rnb_contract_grouped routes scalar-like and tiny contractions to SIMT kernels (csrc/skinny.cu).
    Parity tests on small shapes run twice: as routed, and with the tensor-core kernels forced
    (RNB_SIMT=0; conftest sets RNB_ENV_NOCACHE so the library re-reads its knobs at every call)."""
    if request.param == "tensor":
        monkeypatch.setenv("RNB_SIMT", "0")
    else:
        monkeypatch.delenv("RNB_SIMT", raising=False)
    return request.param
