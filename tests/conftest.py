import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--emulated", action="store_true", default=False,
                     help="run the -m gpu tests against tests/host/_build/libvlo_emul.so (the library compiled for the CPU "
                          "against the SIMT emulator) instead of libvlo.so -- slow, for checking kernel changes without a GPU")


@pytest.fixture(scope="session", autouse=True)
def _emulated_library(request):
    """--emulated: api.Handle talks to the CPU-emulated library for the whole session (TEST INFRASTRUCTURE; the product
    never loads it)."""
    if not request.config.getoption("--emulated"):
        yield
        return
    sys.path.insert(0, os.path.join(ROOT, "tests", "host"))
    import build_emul
    from vil_sensor_fusion_b200 import _lib
    path = build_emul.build()
    if path is None:
        pytest.skip("CUDA headers not found")
    saved = _lib._lib
    _lib._lib = _lib.load(path)
    yield
    _lib._lib = saved


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle
