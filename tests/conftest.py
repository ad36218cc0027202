import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Make sure libhs_b200.so exists (builds it with nvcc when stale; no GPU needed)."""
    sys.path.insert(0, os.path.join(REPO, "multi-uav-pursuit-evasion_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_hs_build", os.path.join(REPO, "multi-uav-pursuit-evasion_b200", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build()
