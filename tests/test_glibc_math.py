"""The device math (hspsquared_b200/csrc/glibc_math_core.h + glibc_tables.h) is a restatement of the libm the reference
reaches through Numba: compiled here for the host with explicit fma() and NO contraction, it must return
the very bits of this box's pow / exp / exp2 / log on millions of arguments.  (The CUDA build runs the same header with
__fma_rn, so the GPU path then equals the oracle bit for bit -- asserted in tests/test_gpu_parity.py.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hspsquared_b200", "csrc")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("glm") / "glibc_math_check")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-I", CSRC, os.path.join(ROOT, "tests", "glibc_math_check.c"),
                    "-o", exe, "-lm"], check=True)
    return exe


def _fma_host():
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((line for line in f if line.startswith("flags")), "") + " "
        return " fma " in flags and " avx2 " in flags
    except OSError:
        return False


@pytest.mark.skipif(not _fma_host(), reason="this host's libm runs glibc's non-FMA builds; the restatement is of the FMA builds")
def test_restatement_equals_this_libm_bit_for_bit(checker):
    r = subprocess.run([checker, "6000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split()[1].startswith("0/") and "log 0/" in r.stdout


def test_generated_tables_match_the_library(tmp_path):
    """The committed header is what tools/extract_glibc_tables.py reads out of this image's libm.so.6."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("extract", os.path.join(ROOT, "tools", "extract_glibc_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.LIBM):
        pytest.skip("no libm at %s" % mod.LIBM)
    out = str(tmp_path / "glibc_tables.h")
    mod.main(out)
    assert open(out).read() == open(mod.OUT).read()


def test_division_by_small_odd_part_constants_equals_ieee_division(tmp_path):
    """csrc/const_div.h: a / 24, a / 100, a / 144 ... in three operations (the header carries the proof); pinned here against the
    host's own division on 12 M arguments per constant -- any bit pattern, exact multiples and their neighbours, short decimals."""
    exe = str(tmp_path / "const_div_check")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-I", CSRC, os.path.join(ROOT, "tests", "const_div_check.c"),
                    "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe, "12000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert all(line.split()[1].startswith("0/") for line in r.stdout.strip().splitlines())
