import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    L = oracle.load()
    L.nso_set_threads(min(8, os.cpu_count() or 1))
    return L


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  Built here (nvcc cross-compiles) if the .so is missing; never skipped silently."""
    from nextsimdg_b200 import capi

    if not os.path.exists(capi.library_path()):
        import __graft_entry__ as g

        g.build_cuda()
    return capi.load_library()
