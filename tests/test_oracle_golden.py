"""Pin the CPU oracle (oracle/hsp2_oracle.c + oracle/wrappers.py).

1. bit-for-bit against the UNMODIFIED reference kernels: sha256 of every float64 series the reference leaves in
   `ts` and every error-count vector, for all cases of oracle/cases.py (tests/golden/reference_hashes.json,
   produced by oracle/gen_golden.py in the build container);
2. against the HSPF known-answer files the reference's own tests use: tests/ipwater/data/Flow_Test.plt with the
   reference's tolerance (tests/ipwater/test_ipwater.py:61,82-84) and tests/test10/HSPFresults/test10I.hbn.
"""
import copy
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import cases as CASES
from oracle import wrappers as W

ACTS = {"ATEMP": W.atemp, "SNOW": W.snow, "PWATER": W.pwater, "IWATER": W.iwater, "SEDMNT": W.sedmnt, "SOLIDS": W.solids, "IWTGAS": W.iwtgas,
        "PSTEMP": W.pstemp, "PWTGAS": W.pwtgas, "PQUAL": W.pqual, "IQUAL": W.iqual}
ALL = CASES.build_cases()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def hashes(golden_dir):
    with open(os.path.join(golden_dir, "reference_hashes.json")) as f:
        return json.load(f)


def run_oracle(case, time_unit):
    si = CASES.siminfo_for(case)
    si["time_unit"] = time_unit
    ts = dict(CASES.forcings(case))
    errs = CASES.run_segment(case["operation"], copy.deepcopy(case["tables"]), ts, si, ACTS)
    return ts, errs


@pytest.mark.parametrize("name", sorted(ALL))
def test_oracle_bit_exact_vs_reference(name, hashes, oracle_lib, golden_dir):
    unit = str(np.load(os.path.join(golden_dir, "calendar_ref.npz"))["time_unit"])
    ts, errs = run_oracle(ALL[name], unit)
    ref = hashes[name]
    assert {a: [int(x) for x in e] for a, e in errs.items()} == ref["errors"]
    for k, (n, h) in ref["series"].items():
        assert k in ts, "oracle did not produce ts[%r]" % k
        assert ts[k].shape[0] == n, k
        assert sha(ts[k]) == h, "series %s differs from the reference" % k


def test_flowtest_hspf_kat(oracle_lib, golden_dir):
    """The reference's own assertions (test_ipwater.py): all error counts 0 and |sum(HSP2 - HSPF)| < 1e-3."""
    hspf = np.load(os.path.join(golden_dir, "hspf_flowtest.npz"))
    ts, errs = run_oracle(ALL["flowtest_pwater"], "us")
    assert not errs["PWATER"].any()
    for v in ("SURO", "IFWO", "AGWO"):
        assert abs((ts[v] - hspf["PERLND1_" + v]).sum()) < 1e-3
        assert np.abs(ts[v] - hspf["PERLND1_" + v]).max() < 1e-6
    ts, errs = run_oracle(ALL["flowtest_iwater"], "us")
    assert not errs["IWATER"].any()
    assert abs((ts["SURO"] - hspf["IMPLND1_SURO"]).sum()) < 1e-3


def test_test10_implnd_hspf_kat(oracle_lib, golden_dir):
    """HSPF's own binary output for test10 IMPLND 1 (fp32): SNOW -> IWATER through the oracle."""
    hspf = np.load(os.path.join(golden_dir, "hspf_test10I.npz"))
    ts, _ = run_oracle(ALL["i001_test10"], "us")
    tol = {"SUPY": 2e-6, "SURO": 2e-6, "SURS": 1e-6, "RETS": 1e-6, "PET": 1e-7, "IMPEV": 1e-7, "PETADJ": 1e-5}
    for v, t in tol.items():
        assert hspf[v].shape[0] == 8784
        assert np.abs(ts[v] - hspf[v]).max() < t, v
    assert abs(ts["SURO"].sum() - hspf["SURO"].sum()) < 1e-4


def test_full_reference_arrays_match(oracle_lib, golden_dir):
    """The committed full fp64 reference arrays agree with the oracle exactly (NaN == NaN)."""
    for name in ("p001_test10", "i001_test10", "cbp_fullchain", "p001_snow_pwater"):
        ref = np.load(os.path.join(golden_dir, "ref_%s.npz" % name))
        ts, _ = run_oracle(ALL[name], "us")
        for k in ref.files:
            assert np.array_equal(ref[k], ts[k], equal_nan=True), (name, k)


def test_test10_implnd_solids_iwtgas_hspf_kat(oracle_lib, golden_dir):
    """HSPF's own hourly SOLIDS and IWTGAS output for test10 IMPLND 1 (test10I.hbn) against the oracle chain
    SNOW -> IWATER -> SOLIDS -> IWTGAS: -1e30 "no runoff" sentinels at the same intervals, values to HSPF's fp32 precision
    (2e-6 of each series' scale)."""
    h = np.load(os.path.join(golden_dir, "hspf_test10I.npz"))
    ts, _ = run_oracle(ALL["i001_iwtgas"], "us")
    for key in ("SOLIDS_SLDS", "SOLIDS_SOSLD", "IWTGAS_SOTMP", "IWTGAS_SODOX", "IWTGAS_SOCO2", "IWTGAS_SOHT",
                "IWTGAS_SODOXM", "IWTGAS_SOCO2M"):
        ref, got = h[key], ts[key.split("_", 1)[1]]
        ok = ref > -1e29
        assert np.array_equal(ok, got > -1e29), key
        assert np.abs(got[ok] - ref[ok]).max() <= 2e-6 * np.abs(ref[ok]).max(), key


def test_test10_implnd_iqual_hspf_kat(oracle_lib, golden_dir):
    """HSPF's own hourly IQUAL output for test10 IMPLND 1 (constituent COD, test10I.hbn) against the oracle chain
    SNOW -> IWATER -> SOLIDS -> IQUAL: storages and fluxes to 2e-6 of scale, the concentrations (a flux divided by a fp32 runoff
    and HSPF's own ft3 conversion constant) to 1e-5."""
    h = np.load(os.path.join(golden_dir, "hspf_test10I.npz"))
    ts, errs = run_oracle(ALL["i001_iqual"], "us")
    assert not errs["IQUAL"].any()
    for cons in ("SQO", "SOQSP", "SOQOC", "SOQC", "SOQS", "SOQO", "SOQUAL"):
        ref, got = h["IQUAL_" + cons], ts["IQUAL1_" + cons]
        ok = ref > -1e29
        assert np.array_equal(ok, got > -1e29), cons
        assert np.abs(got[ok] - ref[ok]).max() <= (1e-5 if cons in ("SOQOC", "SOQC") else 2e-6) * np.abs(ref[ok]).max(), cons
    assert abs(ts["IQUAL1_SOQUAL"].sum() - h["IQUAL_SOQUAL"].sum()) < 1e-5 * h["IQUAL_SOQUAL"].sum()
