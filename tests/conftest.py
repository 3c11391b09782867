import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    need = [os.path.join(g.LIB_DIR, "libahsoka_b200.so"), os.path.join(g.LIB_DIR, "libahsoka_synth.so"),
            os.path.join(ROOT, "oracle", "_ref", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        g.build()
