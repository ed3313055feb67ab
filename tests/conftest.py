import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libmm_b200.so when nvcc is present (CPU container); on the GPU box the prebuilt
    in-tree .so travels with the snapshot and is used as is."""
    from models_b200 import _cabi

    if not _cabi.LIB_PATH.exists():
        from models_b200.csrc import build

        build.build()
    yield


@pytest.fixture()
def device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda", 0)
