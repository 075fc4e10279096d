import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture
def port():
    # reference tests/test_basic.py:18-20
    return random.randint(10000, 50000)


@pytest.fixture(scope="session")
def sim_api():
    """API classes bound to the host-logic simulator (CPU, test-only)."""
    from tests.hostsim import load

    api = load()
    yield api
    api.shutdown()


@pytest.fixture(scope="session")
def cuda_api():
    """API classes bound to the product library (CUDA, sm_100a)."""
    import types

    import starway_b200 as sw

    api = types.SimpleNamespace(
        Server=sw.Server,
        Client=sw.Client,
        ServerEndpoint=sw.ServerEndpoint,
        get_context=sw.get_context,
        shutdown=sw.shutdown,
        backend_name=sw.backend_name,
    )
    yield api
    sw.shutdown()
