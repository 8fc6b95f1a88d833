import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """-m gpu tests need a B200: on a box without a CUDA device they SKIP (they never fall back to
    a CPU path -- there is none -- and they do not error)."""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the consensus path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """The in-tree libraries (oracle always; the CUDA library when it has been built)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def gpu_ctx(built):
    from mandalorion_b200 import PoaContext
    ctx = PoaContext(0)          # no fallback: fails loudly without the CUDA library / GPU
    yield ctx
    ctx.close()
