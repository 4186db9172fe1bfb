import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from refblis import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The real reference library, when oracle/_ref/libblis_ref.so is present."""
    from refblis import RefBlis, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libblis_ref.so not built")
    return RefBlis(threads=min(8, os.cpu_count() or 1))


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through its C ABI.  No fallback: a missing library or a
    missing GPU is an error for -m gpu tests."""
    import torch
    from blis_b200 import _lib, api
    lib = _lib.load()
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.check(lib.b200_init(0), "b200_init")
    return api
