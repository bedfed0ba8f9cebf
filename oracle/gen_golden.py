#!/usr/bin/env python
"""Generate the committed fixtures from the UNMODIFIED reference.  Runs in the build container only
(/root/reference is not shipped to the GPU box); the outputs travel with the repo.

    python -m oracle.gen_golden            (from the repo root)

Writes
  hspsquared_b200/data/test10_met_1976.npz   1976 hourly forcings of test10's PERLND 1 / IMPLND 1, built with the
                                             reference's WDM walkers + get_timeseries/transform conventions, and
                                             the PREC/PETINP columns of tests/ipwater/data/Flow_Test.plt
  tests/golden/reference_hashes.json         sha256 of every float64 series the reference leaves in `ts`, and the
                                             error-count vectors, for every case of oracle/cases.py
  tests/golden/ref_<case>.npz                full float64 reference outputs for a few cases (1e-9 parity targets)
  tests/golden/hspf_flowtest.npz             HSPF PLTGEN columns of Flow_Test.plt (reference test_ipwater KAT)
  tests/golden/hspf_test10I.npz              HSPF binary output test10I.hbn, IMPLND IWATER hourly (KAT)
  tests/golden/calendar_ref.npz              series produced by the reference's pandas calendar code
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cases as CASES  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "hspsquared_b200", "data")
FULL = ("p001_test10", "i001_test10", "cbp_fullchain", "p001_snow_pwater", "i001_iwtgas")


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def gen_met():
    out = {}
    for op, pre in (("PERLND", "P_"), ("IMPLND", "I_")):
        ts, si = H.test10_forcings(op)
        assert si["steps"] == 8784
        for k, v in ts.items():
            out[pre + k] = v
    import pandas as pd

    plt = os.path.join(H.REF_ROOT, "tests", "ipwater", "data", "Flow_Test.plt")
    cols = ["REACH01", "PERLND1_SURO", "PERLND1_IFWO", "PERLND1_AGWO", "IMPLND1_SURO", "PREC", "PETINP"]
    tab = pd.read_table(plt, sep=r"\s+", skiprows=26, names=range(13)).rename(
        columns={i + 6: c for i, c in enumerate(cols)})  # tests/ipwater/test_ipwater.py:23-40
    out["F_PREC"] = tab["PREC"].values.astype(np.float64)
    out["F_PETINP"] = tab["PETINP"].values.astype(np.float64)
    os.makedirs(DATA, exist_ok=True)
    np.savez_compressed(os.path.join(DATA, "test10_met_1976.npz"), **out)
    np.savez_compressed(os.path.join(GOLD, "hspf_flowtest.npz"),
                        **{c: tab[c].values.astype(np.float64) for c in cols[1:5]})


def gen_hbn():
    sys.path.insert(0, H.REF_SRC)
    H.load()
    from hsp2.hsp2tools.HBNOutput import HBNOutput

    hbn = HBNOutput(os.path.join(H.REF_ROOT, "tests", "test10", "HSPFresults", "test10I.hbn"))
    hbn.read_data()
    out = {}
    for cons in ("RETS", "SURS", "PETADJ", "SUPY", "SURO", "PET", "IMPEV"):
        s = hbn.get_time_series("IMPLND", 1, cons, "IWATER", "hourly")
        out[cons] = s.values.astype(np.float64)
        out["_first_stamp"] = np.array(str(s.index[0]))
    # SOLIDS and IWTGAS groups of the same HSPF run (keys prefixed with the section)
    for act, names in (("SOLIDS", ("SLDS", "SOSLD")), ("IWTGAS", ("SOTMP", "SODOX", "SOCO2", "SOHT", "SODOXM", "SOCO2M"))):
        for cons in names:
            out["%s_%s" % (act, cons)] = hbn.get_time_series("IMPLND", 1, cons, act, "hourly").values.astype(np.float64)
    for cons in ("SQO", "SOQSP", "SOQOC", "SOQC", "SOQS", "SOQO", "SOQUAL"):  # IQUAL group; HSPF appends the QUALID (COD)
        out["IQUAL_" + cons] = hbn.get_time_series("IMPLND", 1, cons + "COD", "IQUAL", "hourly").values.astype(np.float64)
    np.savez_compressed(os.path.join(GOLD, "hspf_test10I.npz"), **out)


def ref_activities():
    m = H.load()
    return {"ATEMP": m["ATEMP"].atemp, "SNOW": m["SNOW"].snow, "PWATER": m["PWATER"].pwater,
            "IWATER": m["IWATER"].iwater, "SEDMNT": m["SEDMNT"].sedmnt, "PSTEMP": m["PSTEMP"].pstemp,
            "PWTGAS": m["PWTGAS"].pwtgas, "SOLIDS": m["SOLIDS"].solids, "IWTGAS": m["IWTGAS"].iwtgas,
            "PQUAL": m["PQUAL"].pqual, "IQUAL": m["IQUAL"].iqual}


def run_reference(case):
    import pandas as pd
    from numba import types
    from numba.typed import Dict

    si = CASES.siminfo_for(case)
    si["start"], si["stop"] = pd.Timestamp(si["start"]), pd.Timestamp(si["stop"])
    ts = Dict.empty(key_type=types.unicode_type, value_type=types.float64[:])
    for k, v in CASES.forcings(case).items():
        ts[k] = v
    import copy

    tables = copy.deepcopy(case["tables"])
    errs = CASES.run_segment(case["operation"], tables, ts, si, ref_activities())
    return {k: np.asarray(ts[k]) for k in ts.keys()}, errs


def gen_cases(only=None):
    """only: iterable of case names to (re)generate; the other records of reference_hashes.json are kept."""
    hashes = {}
    path = os.path.join(GOLD, "reference_hashes.json")
    if only and os.path.exists(path):
        hashes = json.load(open(path))
    for name, case in CASES.build_cases().items():
        if only and name not in only:
            continue
        ts, errs = run_reference(case)
        inputs = set(CASES.forcings(case).keys())
        rec = {"series": {}, "errors": {a: [int(x) for x in e] for a, e in errs.items()}, "sums": {}}
        for k, v in ts.items():
            if k in inputs and k not in ("AIRTMP",):
                continue
            if k == "SVP":
                continue
            rec["series"][k] = [int(v.shape[0]), sha(v)]
            rec["sums"][k] = float(np.nansum(np.where(np.abs(v) > 1e29, 0.0, v)))
        hashes[name] = rec
        print("%-22s %3d series  errors=%s" % (name, len(rec["series"]), rec["errors"]))
        if name in FULL:
            keep = {k: v for k, v in ts.items() if k in rec["series"] and v.shape[0] == case["steps"]}
            np.savez_compressed(os.path.join(GOLD, "ref_%s.npz" % name), **keep)
    with open(os.path.join(GOLD, "reference_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)


def gen_calendar():
    m = H.load()
    U, S = m["utilities"], m["SNOW"]
    tab = [0.04, 0.04, 0.03, 0.03, 0.03, 0.03, 0.10, 0.17, 0.19, 0.14, 0.05, 0.04]
    out = {}
    spans = [("1976-01-01", "1977-01-01", 60), ("1976-01-01", "1977-01-01", 30), ("1976-01-01", "1977-01-01", 120),
             ("1976-03-15", "1978-07-01", 60), ("1984-01-01", "1986-01-01", 15), ("1976-06-01 06:00", "1976-09-01", 60),
             ("1976-01-01", "1979-01-01", 1440), ("1976-01-01", "1977-01-01", 360)]
    for i, (a, b, d) in enumerate(spans):
        si = H.make_siminfo(a, b, d)
        out["%d_HRFG" % i] = U.hoursval(si, np.ones(24), dofirst=True)
        out["%d_DAYFG" % i] = U.hourflag(si, 0, dofirst=True)
        out["%d_HR1FG" % i] = U.hourflag(si, 1, dofirst=True)
        out["%d_HR6IND" % i] = S.hour6flag(si, dofirst=True)
        out["%d_LAPSE" % i] = U.hoursval(si, U.LAPSE, lapselike=True)
        out["%d_SEASONS" % i] = U.monthval(si, U.SEASONS).astype(float)
        out["%d_DAYVAL" % i] = U.dayval(si, tab)
    out["spans"] = np.array(json.dumps(spans))
    out["table"] = np.array(tab)
    import pandas as pd

    out["time_unit"] = np.array(str(pd.date_range("2000-01-01", periods=2, freq="MS").unit))
    np.savez_compressed(os.path.join(GOLD, "calendar_ref.npz"), **out)


if __name__ == "__main__":
    if not H.available():
        sys.exit("reference tree not found at %s" % H.REF_ROOT)
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1:  # python -m oracle.gen_golden <case> ...: only those cases (and the HSPF file)
        gen_hbn()
        gen_cases(set(sys.argv[1:]))
        sys.exit(0)
    gen_met()
    gen_hbn()
    gen_calendar()
    gen_cases()
