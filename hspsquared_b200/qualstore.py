"""Host image of a PQUAL / IQUAL batch: one ROW per (segment, constituent).

The reference runs the constituents of a segment in a loop inside `_pqual_` / `_iqual_` with per-constituent keys
(`ui['QSDFG1']`, `ts['POTFW1']`, `ts['PQUAL1_SQO']` ..., PQUAL.py:52-96, 148-232; IQUAL.py:48-94, 131-183); constituents
never interact, so the batch flattens (segment, constituent) pairs into rows of an ordinary SegmentStore and the kernels
(csrc/qual.cuh) treat a row like a land segment whose forcings are the segment's flow / sediment / solids outflow series.

`uci` below is `uci[(operation, 'PQUAL' | 'IQUAL', segment)]` as HDF5.read_uci builds it (hsp2io/hdf.py:34-68, from readUCI.py:
690-711): 'PARAMETERS' (NQUAL + LAT-FACTOR), optional 'FLAGS' (PQADFG1.. / IQADFG1..), and per constituent k
'<X>QUALk_FLAGS', '<X>QUALk_PARAMETERS', '<X>QUALk_MONTHLY/<NAME>'.
"""
import numpy as np

from . import _lib
from .segstore import SegmentStore, _initm, _table, make_ui

PQUAL_INPUTS = ("SURO", "IFWO", "AGWO", "PERO", "WSSD", "SCRSD", "PREC", "SLIQSP", "ILIQC", "ALIQC")
IQUAL_INPUTS = ("SURO", "SOSLD", "PREC", "SLIQSP")


def nquals(uci):
    """PQUAL.py:34-38"""
    if "PARAMETERS" in uci and "NQUAL" in uci["PARAMETERS"]:
        return int(uci["PARAMETERS"]["NQUAL"])
    return 1


def constituent(operation, uci, k):
    """Effective values of constituent k (1-based): {'raw', 'flag', 'mon', 'sqo', 'dep_series', 'dep_without_qsofg'}."""
    imp = operation == "IMPLND"
    pre = ("IQUAL" if imp else "PQUAL") + str(k)
    f, p = uci[pre + "_FLAGS"], uci[pre + "_PARAMETERS"]
    f["QUALID"]  # the reference indexes it (PQUAL.py:41)
    ui = make_ui(uci)
    raw, flag, mon = {}, {}, {}
    qsofg = f["QSOFG"]
    raw["Q_WSFAC"] = 2.30 / p["WSQOP"]
    raw["SLIFAC"] = ui["SLIFAC"]
    if not imp:
        ui["SDLFAC"]  # read by the reference (PQUAL.py:128)
        raw["ILIFAC"], raw["ALIFAC"] = ui["ILIFAC"], ui["ALIFAC"]
    flag["Q_QSDFG"] = int(bool(f["QSDFG"]))
    flag["Q_QSO_NZ"] = int(qsofg != 0)
    flag["Q_QSO_GE1"] = int(qsofg >= 1)
    flag["Q_QSO_EQ1"] = int(qsofg == 1)
    flag["Q_QSO_EQ2"] = int(qsofg == 2)
    flag["Q_QSO_M1"] = int(qsofg == -1)
    if imp and qsofg not in (-1, 0) and not qsofg >= 1:
        raise NotImplementedError("IQUAL QSOFG=%r: neither branch of IQUAL.py:224-269 runs and SOQO is whatever was left" % qsofg)
    tables = [("POTFW", "VPFWFG", "POTFW"), ("ACQOP", "VQOFG", "ACQOP")]
    if not imp:
        tables += [("POTFS", "VPFSFG", "POTFS"), ("IOQC", "VIQCFG", "IOQC"), ("AOQC", "VAQCFG", "AOQC")]
        flag["Q_QIFWFG"], flag["Q_QAGWFG"] = int(f["QIFWFG"] != 0), int(bool(f["QAGWFG"]))
        flag["Q_VIQC34"], flag["Q_VAQC34"] = int(f["VIQCFG"] in (3, 4)), int(f["VAQCFG"] in (3, 4))
    for name, vflag, tab in tables:
        raw["Q_" + name] = float(p[name])
        if _initm(uci, f[vflag], pre + "_MONTHLY/" + tab):
            flag["M_Q_" + name] = 1
            mon["Q_" + name] = _table(uci, pre + "_MONTHLY/" + tab)
    p["SQOLIM"]  # initm(... SQOLIM) is evaluated even though the kernels never read the series (PQUAL.py:76)
    # REMQOP = ACQOP / SQOLIM, month by month when VQOFG and the SQOLIM table exist (utilities.py:214-225)
    if f["VQOFG"] and (pre + "_MONTHLY/SQOLIM") in uci:
        a, b = _table(uci, pre + "_MONTHLY/ACQOP"), _table(uci, pre + "_MONTHLY/SQOLIM")
        flag["M_Q_REMQOP"] = 1
        mon["Q_REMQOP"] = [a[m] / b[m] for m in range(12)]
        raw["Q_REMQOP"] = 0.0
    else:
        raw["Q_REMQOP"] = p["ACQOP"] / p["SQOLIM"]
    # atmospheric deposition (PQUAL.py:80-96): flag > 0 monthly table (else 0), -1 an input series, 0 none
    dep_series = {}
    adf = adc = 0
    if "FLAGS" in uci:
        u = uci["FLAGS"]
        adname = "IQAD" if imp else "PQAD"
        adf, adc = u[adname + "FG" + str(2 * k - 1)], u[adname + "FG" + str(2 * k)]
        for val, nm, col in ((adf, "FX", "Q_ADFX"), (adc, "CN", "Q_ADCN")):
            if val > 0:
                if _initm(uci, val, pre + "_MONTHLY/" + adname + nm):
                    flag["M_" + col] = 1
                    mon[col] = _table(uci, pre + "_MONTHLY/" + adname + nm)
            elif val == -1:
                dep_series[col[2:]] = adname + nm + str(k) + " 1"  # ts key the reference reads
    raw["Q_ADFX"] = raw["Q_ADCN"] = 0.0
    sqo = float(p["SQO"]) if (imp or qsofg) else 0.0  # PQUAL.py:161-164
    return {"raw": raw, "flag": flag, "mon": mon, "sqo": sqo, "dep_series": dep_series,
            "dep_without_qsofg": int(qsofg == 0 and (adf != 0 or adc != 0))}


def qual_store(operation, uci_list, calendar):
    """SegmentStore of all constituents of all segments (`uci_list`: one PQUAL / IQUAL table set per segment).
    Extra attributes: rows [(segment index, k)], dep_without_qsofg int64[rows] (the reference's errors[0] increments),
    dep_series [rows] {'ADFX' | 'ADCN': ts key} for constituents whose deposition is an input series."""
    act = _lib.ACT_BITS["IQUAL" if operation == "IMPLND" else "PQUAL"]
    rows, effs = [], []
    for i, uci in enumerate(uci_list):
        for k in range(1, nquals(uci) + 1):
            rows.append((i, k))
            effs.append(constituent(operation, uci, k))
    n = len(rows)
    st = SegmentStore(operation, act, n, float(calendar.delt))
    for j, e in enumerate(effs):
        for name, v in e["raw"].items():
            st.par[st.P[name], j] = v
        for name, v in e["flag"].items():
            if v:
                st.flags[j] |= np.uint64(1) << np.uint64(st.FB[name])
        st.state[st.S["Q_SQO"], j] = e["sqo"]
    for name in sorted({m for e in effs for m in e["mon"]}):
        tab = np.zeros((12, n))
        for j, e in enumerate(effs):
            if name in e["mon"]:
                tab[:, j] = e["mon"][name]
        st.monthly[st.M[name]] = tab
    st.rows = rows
    st.dep_without_qsofg = np.array([e["dep_without_qsofg"] for e in effs], dtype=np.int64)
    st.dep_series = [e["dep_series"] for e in effs]
    return st
