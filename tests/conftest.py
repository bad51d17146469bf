import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running known-answer test")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle_lib
    oracle_lib.build()
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def b2s():
    """The product package (CUDA library behind the C ABI). Fails loudly if the library is not built."""
    import b200stencil  # noqa: F401  (registers the package)
    from b200stencil import capi
    capi.lib()
    return b200stencil


@pytest.fixture(scope="session")
def gpu(b2s):
    from b200stencil import capi
    n = capi.device_count()
    assert n > 0, "no CUDA device: -m gpu tests must run on the GPU box (there is no CPU fallback)"
    return n
