import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def oracle_lib():
    """Build (gcc, seconds) and load the CPU oracle."""
    from oracle import oracle

    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="module")
def emu_library(oracle_lib):
    """Point the loader at the host-compiled TEST DOUBLE of libhsp2g for this module (tests/emu/README.md)."""
    from hspsquared_b200 import _lib
    from tests.emu import build_emu

    path = build_emu.build()
    saved = (_lib.LIB_PATH, _lib._lib, dict(_lib._names), set(_lib.ALLOW_BUILDS))
    _lib.LIB_PATH, _lib._lib = path, None
    _lib.ALLOW_BUILDS.add("host-emu")  # only the CPU test suite admits the double
    _lib._names.clear()
    yield path
    _lib.LIB_PATH, _lib._lib = saved[0], saved[1]
    _lib.ALLOW_BUILDS.clear()
    _lib.ALLOW_BUILDS.update(saved[3])
    _lib._names.clear()
    _lib._names.update(saved[2])
