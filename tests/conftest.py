import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def lz():
    import lanczos_hls_b200
    from lanczos_hls_b200 import build
    if not os.path.exists(lanczos_hls_b200.lib_path()):
        build.build()
    lanczos_hls_b200.lib()
    return lanczos_hls_b200
