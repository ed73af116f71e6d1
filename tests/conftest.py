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


@pytest.fixture
def knobs(ctx):
    """Tuning knobs of the session context (rtw_cuda_set_option; the library reads the environment only once, at
    rtw_cuda_create).  knobs.setenv(name, value) mirrors monkeypatch.setenv; everything is restored afterwards."""
    touched = []

    class K:
        @staticmethod
        def setenv(name, value):
            ctx.set_option(name, value)
            touched.append(name)

    yield K
    for name in touched:
        ctx.set_option(name, None)
