import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "blocked_tensordot.npz"))


@pytest.fixture(params=["auto", "tensor"])
def contract_route(request, monkeypatch):
    """rnb_contract_grouped routes scalar-like and tiny contractions to SIMT kernels (csrc/skinny.cu).
    Parity tests on small shapes run twice: as routed, and with the tensor-core kernels forced
    (RNB_SIMT=0, read by the library at every call)."""
    if request.param == "tensor":
        monkeypatch.setenv("RNB_SIMT", "0")
    else:
        monkeypatch.delenv("RNB_SIMT", raising=False)
    return request.param
