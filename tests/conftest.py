import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def ctx():
    """One mimosa_b200 context on cuda:0 for the whole GPU session.  Fails loudly without a GPU."""
    import mimosa_b200

    c = mimosa_b200.Context(0)
    yield c
    c.close()
