import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def azb():
    """The product binding (alphazero-rs_b200/__init__.py over libazb200.so)."""
    import __graft_entry__ as ge
    ge.build_product()
    return importlib.import_module("alphazero-rs_b200")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/, test infrastructure only)."""
    import __graft_entry__ as ge
    ge.build_oracle()
    import oracle_api
    return oracle_api
