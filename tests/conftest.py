import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rtw():
    import rtw_b200
    return rtw_b200


@pytest.fixture(scope="session")
def oracle():
    import rtw_b200  # noqa: F401  (abi structs)
    import oracle_binding
    oracle_binding.lib()
    return oracle_binding


@pytest.fixture(scope="session")
def earth_rgba():
    import rtw_b200
    return rtw_b200.host_lib.decode_png(rtw_b200.host_lib.ASSET_EARTH)


@pytest.fixture(scope="session")
def ctx(rtw):
    """One CUDA context for the whole GPU session.  Raises (never skips, never falls back) without a GPU."""
    c = rtw.Context(0)
    yield c
    c.close()
