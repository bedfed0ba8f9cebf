"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/hsp2g.h declares; without a GPU
the product path fails loudly instead of falling back (no compute calls are made here)."""
import ctypes as C
import os
import re

import pytest

from hspsquared_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()  # nvcc cross-compiles sm_100a without a GPU
    assert _lib.LIB_PATH == os.path.join(ROOT, "hspsquared_b200", "libhsp2g.so")
    return _lib.load()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "hsp2g.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(hsp2g_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libhsp2g.so does not export %s" % s
    assert set(syms) == set(_lib.EXPORTS), "ctypes prototypes and the header disagree"
    assert lib.hsp2g_abi_version() == _lib.ABI_VERSION == 3
    assert _lib.build_kind() == "cuda/sm_100a"


def test_name_tables(lib):
    for table, must in (("params", "KGW"), ("monthly", "CEPSC"), ("flags", "M_CEPSC"), ("states", "PDEPTH"),
                        ("forcings", "PETINP"), ("outputs", "PWATER_SURO"), ("errors", "PWATER6")):
        names = _lib.names(table)
        assert must in names and len(set(names)) == len(names)
    assert len(_lib.names("flags")) <= 64
    assert len([n for n in _lib.names("outputs") if n.startswith("SNOW_")]) == 22  # SNOW.py:162-183
    assert len([n for n in _lib.names("outputs") if n.startswith("PWATER_")]) == 31  # PWATER.py:123-153
    assert len([n for n in _lib.names("outputs") if n.startswith("IWATER_")]) == 8  # IWATER.py:119-126
    assert len([n for n in _lib.names("outputs") if n.startswith("PWTGAS_")]) == 21  # PWTGAS.py:143-163


def test_is_sm100a_cuda_code():
    """The product library carries sm_100a device code (cuobjdump), i.e. it is the CUDA build, not a host double."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.hsp2g_batch_create(0, 0, 4, 6, 60.0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.hsp2g_last_error()
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch
    from hspsquared_b200.segstore import SegmentStore
    from hspsquared_b200 import test10

    cal = Calendar("1976-01-01", "1976-01-03", 60, "us")
    t = test10.perlnd()
    t["ACTIVITY"].update({"PSTEMP": 0, "PWTGAS": 0})
    store = SegmentStore.from_tables("PERLND", [t], cal)
    with pytest.raises(_lib.Hsp2gError):
        LandBatch(store, cal, ["PREC", "PETINP", "AIRTMP", "WINMOV", "SOLRAD", "DTMPG"], "all")


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.Hsp2gError):
        _lib.load()


def test_loader_refuses_a_host_build(emu_library, monkeypatch):
    """HSP2G_LIB may name another build of the library, but the product loader opens CUDA builds only: the host-compiled test
    double (the one thing in this repo that could pass for a CPU path) is refused unless the test suite itself admits it."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "ALLOW_BUILDS", {_lib.CUDA_BUILD})
    with pytest.raises(_lib.Hsp2gError, match="host-emu"):
        _lib.load()
