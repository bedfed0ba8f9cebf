"""Drop-in activities behind HSP2's module contract:  function(io_manager, siminfo, uci, ts) -> (errors, errmsgs).

`install()` swaps these into the reference's registry `hsp2.hsp2.configuration.activities` (configuration.py:45-55 --
the same dict object main.py:15 and mainDoE.py:16 imported), so `hsp2.main` dispatches land activities to libhsp2g
without any reference file being edited.  Each call has the reference's granularity -- ONE segment,
ONE section (main.py:249-250): inputs are read from `ts`, every output series the reference would create is written back
into `ts` as float64[steps], the error-count vector and message tuple have the reference's shapes, and the two dict side
effects are kept (SNOW.py:60-63 siminfo['ICEFG'], PWATER.py:89-93 uci['PARAMETERS']['CSNOFG']).  When the caller is
`hsp2.main`, the work is NOT done at that granularity: the first call of a segment advances a whole window of the
operation sequence's segments through all their sections in one fused batch and the later calls are served from its
columns (lazybatch.py); a call that cannot be matched to such a batch runs its one segment, one section on the GPU by
itself.  Ensembles that do not need the reference dispatcher use hspsquared_b200.engine.LandBatch directly.  There is no
CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .calendar import SEASONS, Calendar
from .engine import LandBatch, output_names, pack_forcings
from .segstore import SegmentStore

ERRMSGS = {
    "ATEMP": (),
    "SNOW": ("Snow simulation cannot function properly with delt> 360",),  # SNOW.py:23-24
    "PWATER": (  # PWATER.py:15-25
        "PWATER: Sum of irrtgt in not one", "PWATER: UZRAA exceeds UZRA array bounds",
        "PWATER: INTGB exceeds INTGRL array bounds", "PWATER: Reduced AGWO value to available",
        "PWATER: Reduced GVWS to AGWS", "PWATER: GWVS < -0.02, set to zero",
        "PWATER: Proute runoff did not converge", "PWATER: UZI highly negative", "PWATER: Reset AGWS to zero",
        "PWATER: High Water Table code not implemented"),
    "IWATER": ("IWATER: IROUTE Newton Method did not converge",),  # IWATER.py:15-16
    "SEDMNT": [],  # SEDMNT.py:11
    "PSTEMP": ["SLTMP temperature less than -100C", "ULTMP temperature less than -100C",  # PSTEMP.py:13-18
               "LGTMP temperature less than -100C", "SLTMP temperature greater than 100C",
               "ULTMP temperature greater than 100C", "LGTMP temperature greater than 100C"],
    "PWTGAS": [],  # PWTGAS.py:12
    "SOLIDS": [],  # SOLIDS.py:14
    "IWTGAS": [],  # IWTGAS.py:11
    "PQUAL": ("PQUAL: A constituent must be associated with overland flow in order to receive atmospheric deposition inputs", ""),
    "IQUAL": ("IQUAL: A constituent must be associated with overland flow in order to receive atmospheric deposition inputs", ""),
}

# series each section may read from ts (kernel forcing ids)
INPUTS = {
    "ATEMP": ("GATMP", "PREC"),
    "SNOW": ("AIRTMP", "PREC", "DTMPG", "WINMOV", "SOLRAD", "CLOUD"),
    "PWATER": ("PETINP", "PREC", "AIRTMP", "RAINF", "SNOCOV", "WYIELD", "PACKI", "SURLI", "UZLI", "IFWLI", "LZLI",
               "AGWLI", "LGTMP"),
    "IWATER": ("PETINP", "PREC", "AIRTMP", "RAINF", "SNOCOV", "WYIELD", "SURLI"),
    "SEDMNT": ("RAINF", "PREC", "SNOCOV", "SURO", "SURS", "SLSED"),
    "PSTEMP": ("AIRTMP",),
    "PWTGAS": ("WYIELD", "SURO", "IFWO", "AGWO", "SURLI", "IFWLI", "AGWLI", "SLTMP", "ULTMP", "LGTMP", "SLITMP",
               "SLIDOX", "SLICO2", "ILITMP", "ILIDOX", "ILICO2", "ALITMP", "ALIDOX", "ALICO2"),
    "SOLIDS": ("SURO", "SURS", "PREC"),
    "IWTGAS": ("AIRTMP", "WYIELD", "SURO", "SURLI", "SLITMP", "SLIDOX", "SLICO2"),
}
DEVICE = 0


def _calendar(siminfo):
    key = (str(siminfo["start"]), str(siminfo["stop"]), int(siminfo["delt"]), siminfo.get("time_unit"))
    cache = siminfo.get("_hsp2g_calendar")
    if cache is None or cache[0] != key:
        cache = (key, Calendar.from_siminfo(siminfo, siminfo.get("time_unit")))
        siminfo["_hsp2g_calendar"] = cache
    return cache[1]


def _run_section(section, operation, siminfo, uci, ts, **eff_kw):
    units = int(siminfo.get("units", 1))  # 2: metric, every section converts where the reference does (segstore.py, csrc/)
    cal = _calendar(siminfo)
    steps = int(siminfo["steps"])
    if steps != cal.steps:
        raise ValueError("siminfo['steps'] does not match start/stop/delt")
    # called from hsp2.main: all segments of the window were advanced together, fused, on the first call (lazybatch.py)
    from . import lazybatch

    served = lazybatch.serve(section, operation, siminfo, uci, ts, cal, DEVICE)
    if served is not None:
        return served, cal
    tables = {"ACTIVITY": {section: 1}, section: uci}
    names = [n for n in INPUTS[section] if n in ts]
    store = SegmentStore.from_tables(operation, [tables], cal, ext_names=[k for k in ts.keys()], units=units, **eff_kw)
    save = output_names(operation, store.act)
    with LandBatch(store, cal, names, save, device=DEVICE) as batch:
        forcing = pack_forcings({n: np.asarray(ts[n]) for n in names}, names, 1)
        out = batch.run(forcing)
        errors = batch.errors_for(section)
    for r, full in enumerate(save):
        ts[full.split("_", 1)[1]] = np.ascontiguousarray(out[:, r, 0])
    return errors, cal


def _full(steps, v):
    return np.full(steps, float(v))


def atemp(io_manager, siminfo, uci, ts):
    """ATEMP.py:16-30"""
    errors, cal = _run_section("ATEMP", siminfo.get("_operation", "PERLND"), siminfo, uci, ts)
    ts["LAPSE"] = cal.lapse()
    return errors, ERRMSGS["ATEMP"]


def snow(io_manager, siminfo, uci, ts):
    """SNOW.py:27-88"""
    if siminfo["delt"] > 360:
        raise KeyError("ICEFLG")  # what the reference does for delt > 360 (SNOW.py:85 typo); mirrored
    steps = siminfo["steps"]
    errors, cal = _run_section("SNOW", siminfo.get("_operation", "PERLND"), siminfo, uci, ts)
    # helper series the wrapper leaves in ts (visible to saveall and to later sections)
    ts["SEASONS"] = cal.monthval(SEASONS)
    for name in ("AIRTMP", "CLOUD", "DTMPG", "PREC", "SOLRAD", "WINMOV"):
        if name not in ts:
            ts[name] = np.full(steps, np.nan)
    u = uci["PARAMETERS"]
    for name in ("CCFACT", "COVIND", "MGMELT", "MWATER", "SHADE", "SNOEVP", "SNOWCF"):
        if name not in ts and name in u:
            ts[name] = _full(steps, u[name])
    ts["HR6IND"] = cal.hour6flag(dofirst=True)
    ts["HRFG"] = cal.hoursval(np.ones(24), dofirst=True)
    vkmfg = (uci.get("FLAGS") or {}).get("VKMFG", 0)
    ts["KMELT"] = cal.initm(uci, vkmfg, "MONTHLY_KMELT", u["KMELT"])
    siminfo["ICEFG"] = 0
    if "FLAGS" in uci and "ICEFG" in uci["FLAGS"]:
        siminfo["ICEFG"] = uci["FLAGS"]["ICEFG"]
    return errors, ERRMSGS["SNOW"]


def pwater(io_manager, siminfo, uci, ts):
    """PWATER.py:27-99"""
    steps = siminfo["steps"]
    icefg = siminfo["ICEFG"] if "ICEFG" in siminfo else 0.0
    errors, cal = _run_section("PWATER", "PERLND", siminfo, uci, ts, pw_icefg=icefg)
    for name in ("AGWLI", "IFWLI", "LGTMP", "LZLI", "PETINP", "PREC", "SURLI", "UZLI"):
        if name not in ts:
            ts[name] = np.zeros(steps)
    for name in ("AIRTMP", "PACKI", "PETINP", "PREC", "RAINF", "SNOCOV", "WYIELD"):
        if name not in ts:
            ts[name] = np.full(steps, np.nan)
    u = uci["PARAMETERS"]
    for name in ("AGWRC", "DEEPFR", "INFILT", "KVARY", "LZSN", "PETMIN", "PETMAX"):
        if name not in ts and name in u:
            ts[name] = _full(steps, u[name])
    if "VLEFG" in u:
        rules = {"LZETP": (u["VLEFG"] == 1) or (u["VLEFG"] == 3), "CEPSC": u["VCSFG"], "INTFW": u["VIFWFG"],
                 "IRC": u["VIRCFG"], "NSUR": u["VNNFG"], "UZSN": u["VUZFG"]}
        for n, f in rules.items():
            ts[n] = cal.initm(uci, f, "MONTHLY_" + n, u[n])
    else:
        for n in ("LZETP", "CEPSC", "INTFW", "IRC", "NSUR", "UZSN"):
            ts[n] = _full(steps, u[n])
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    ts["HRFG"] = cal.hoursval(np.ones(24), dofirst=True)
    csnofg = u.get("CSNOFG", 0)
    u["CSNOFG"] = int(csnofg) if type(csnofg) in (int, float) else 0  # PWATER.py:89-93
    return errors, ERRMSGS["PWATER"]


def iwater(io_manager, siminfo, uci, ts):
    """IWATER.py:19-72"""
    steps = siminfo["steps"]
    errors, cal = _run_section("IWATER", "IMPLND", siminfo, uci, ts)
    for name in ("AIRTMP", "PETINP", "PREC", "RAINF", "SNOCOV", "WYIELD"):
        if name not in ts:
            ts[name] = np.full(steps, np.nan)
    if "SURLI" not in ts:
        ts["SURLI"] = np.zeros(steps)
    u = uci["PARAMETERS"]
    for name in ("PETMAX", "PETMIN"):
        if name not in ts:
            ts[name] = _full(steps, u[name])
    if "VRSFG" in u:
        ts["RETSC"] = cal.initm(uci, u["VRSFG"], "MONTHLY_RETSC", u["RETSC"])
        ts["NSUR"] = cal.initm(uci, u["VNNFG"], "MONTHLY_NSUR", u["NSUR"])
    else:
        ts["RETSC"] = _full(steps, u["RETSC"])
        ts["NSUR"] = _full(steps, u["NSUR"])
    ts["HR1FG"] = cal.hourflag(1, dofirst=True)
    ts["HRFG"] = cal.hoursval(np.ones(24), dofirst=True)
    return errors, ERRMSGS["IWATER"]


def sedmnt(io_manager, siminfo, uci, ts):
    """SEDMNT.py:21-49"""
    n = siminfo["steps"]
    errors, cal = _run_section("SEDMNT", "PERLND", siminfo, uci, ts)
    u = uci["PARAMETERS"]
    ts["COVERI"] = cal.initm(uci, u["CRVFG"], "MONTHLY_COVER", u["COVER"]) if "CRVFG" in u else _full(n, u["COVER"])
    ts["NVSI"] = cal.initm(uci, u["VSIVFG"], "MONTHLY_NVSI", u["NVSI"]) if "VSIVFG" in u else _full(n, u["NVSI"])
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    for name in ("PREC", "SLSED", "SNOCOV", "SURO", "SURS"):
        if name not in ts:
            ts[name] = np.zeros(n)
    return errors, ERRMSGS["SEDMNT"]


def pstemp(io_manager, siminfo, uci, ts):
    """PSTEMP.py:23-59"""
    n = siminfo["steps"]
    errors, cal = _run_section("PSTEMP", "PERLND", siminfo, uci, ts)
    u = uci["PARAMETERS"]
    for fl, names in (("SLTVFG", ("ASLT", "BSLT")), ("ULTVFG", ("ULTP1", "ULTP2")), ("LGTVFG", ("LGTP1", "LGTP2"))):
        for nm in names:
            ts[nm] = cal.initm(uci, u[fl], "MONTHLY_" + nm, u[nm]) if fl in u else _full(n, u[nm])
    ts["HRFG"] = cal.hoursval(np.ones(24), dofirst=True)
    return errors, ERRMSGS["PSTEMP"]


def pwtgas(io_manager, siminfo, uci, ts):
    """PWTGAS.py:28-62"""
    n = siminfo["steps"]
    errors, cal = _run_section("PWTGAS", "PERLND", siminfo, uci, ts)
    u = uci["PARAMETERS"]
    for fl, nm in (("IDVFG", "IDOXP"), ("ICVFG", "ICO2P"), ("GDVFG", "ADOXP"), ("GCVFG", "ACO2P")):
        ts[nm] = cal.initm(uci, u[fl], "MONTHLY_" + nm, u[nm]) if fl in u else _full(n, u[nm])
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    for name in ("WYIELD", "SURO", "IFWO", "AGWO", "SURLI", "IFWLI", "AGWLI"):
        if name not in ts:
            ts[name] = np.zeros(n)
    for name in ("SLTMP", "ULTMP", "LGTMP", "SLITMP", "SLIDOX", "SLICO2", "ILITMP", "ILIDOX", "ILICO2", "ALITMP",
                 "ALIDOX", "ALICO2"):
        if name not in ts:
            ts[name] = np.full(n, -1.0e30)
    return errors, ERRMSGS["PWTGAS"]


def solids(io_manager, siminfo, uci, ts):
    """SOLIDS.py:16-49"""
    n = siminfo["steps"]
    errors, cal = _run_section("SOLIDS", "IMPLND", siminfo, uci, ts)
    for name in ("SURO", "SURS", "PREC", "SLSLD"):
        if name not in ts:
            ts[name] = np.zeros(n)
    u = uci["PARAMETERS"]
    ts["ACCSDP"] = cal.initm(uci, u["VASDFG"], "MONTHLY_ACCSDP", u["ACCSDP"]) if "VASDFG" in u else _full(n, u["ACCSDP"])
    ts["REMSDP"] = cal.initm(uci, u["VRSDFG"], "MONTHLY_REMSDP", u["REMSDP"]) if "VRSDFG" in u else _full(n, u["REMSDP"])
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    return errors, ERRMSGS["SOLIDS"]


def iwtgas(io_manager, siminfo, uci, ts):
    """IWTGAS.py:33-63.  Like the reference, raises KeyError when the IWT-INIT states are absent (IWTGAS.py:79-81)."""
    n = siminfo["steps"]
    errors, cal = _run_section("IWTGAS", "IMPLND", siminfo, uci, ts)
    u = uci["PARAMETERS"]
    for nm in ("AWTF", "BWTF"):
        ts[nm] = cal.initm(uci, u["WTFVFG"], "MONTHLY_" + nm, u[nm]) if "WTFVFG" in u else _full(n, u[nm])
    for name in ("AIRTMP", "WYIELD", "SURO", "SURLI"):
        if name not in ts:
            ts[name] = np.zeros(n)
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    for name in ("SLITMP", "SLIDOX", "SLICO2"):
        if name not in ts:
            ts[name] = np.full(n, -1.0e30)
    return errors, ERRMSGS["IWTGAS"]


def _qual(operation, siminfo, uci, ts):
    """PQUAL.py:26-113 / IQUAL.py:24-107: all constituents of the segment in one launch, one row each."""
    from .qualstore import IQUAL_INPUTS, PQUAL_INPUTS, nquals, qual_store

    # no unit handling in the reference's PQUAL / IQUAL: a metric run reads the flows in the units they were stored in
    imp = operation == "IMPLND"
    sec, adname = ("IQUAL", "IQAD") if imp else ("PQUAL", "PQAD")
    cal = _calendar(siminfo)
    n = int(siminfo["steps"])
    if n != cal.steps:
        raise ValueError("siminfo['steps'] does not match start/stop/delt")
    nq = nquals(uci)
    store = qual_store(operation, [uci], cal)
    # helper series the wrapper leaves in ts (visible to saveall)
    u = uci.get("FLAGS")
    for k in range(1, nq + 1):
        K = str(k)
        f, p = uci[sec + K + "_FLAGS"], uci[sec + K + "_PARAMETERS"]
        tabs = [("POTFW", "VPFWFG", "POTFW"), ("ACQOP", "VQOFG", "ACQOP"), ("SQOLIM", "VQOFG", "SQOLIM")]
        if not imp:
            tabs[1:1] = [("POTFS", "VPFSFG", "POTFS")]
            tabs += [("IOQCP", "VIQCFG", "IOQC"), ("AOQCP", "VAQCFG", "AOQC")]
        for key, vflag, tab in tabs:
            ts[key + K] = cal.initm(uci, f[vflag], sec + K + "_MONTHLY/" + tab, p[tab])
        ts["REMQOP" + K] = cal.initmdiv(uci, f["VQOFG"], sec + K + "_MONTHLY/ACQOP", sec + K + "_MONTHLY/SQOLIM",
                                        p["ACQOP"], p["SQOLIM"])
        ts[adname + "FX" + K] = np.zeros(n)
        ts[adname + "CN" + K] = np.zeros(n)
        if u is not None:
            for nm, val in (("FX", u[adname + "FG" + str(2 * k - 1)]), ("CN", u[adname + "FG" + str(2 * k)])):
                if val > 0:
                    ts[adname + nm + K] = cal.initm(uci, val, sec + K + "_MONTHLY/" + adname + nm, 0.0)
                elif val == -1:
                    ts[adname + nm + K] = ts[adname + nm + K + " 1"]
    if imp:
        for name in ("SLIQSX", "SLIQO", "SLIQSP"):  # IQUAL.py:96-98
            if name not in ts:
                ts[name] = np.full(n, -1.0e30)
    else:
        for name in ("SLIQSP", "ILIQC", "ALIQC"):  # PQUAL.py:98-100
            if name not in ts:
                ts[name] = np.zeros(n)
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    if not imp:
        for name in ("SURO", "IFWO", "AGWO", "PERO", "WSSD", "SCRSD"):  # PQUAL.py:104-106
            if name not in ts:
                ts[name] = np.zeros(n)
    names = [nm for nm in (IQUAL_INPUTS if imp else PQUAL_INPUTS) if nm in ts]
    forcing = np.empty((n, len(names) + 2, nq))
    for r, nm in enumerate(names):
        forcing[:, r, :] = np.asarray(ts[nm], dtype=np.float64)[:n, None]
    if any(store.dep_series):  # a constituent's deposition is an input series: every row then reads its own column
        for k in range(1, nq + 1):
            forcing[:, len(names), k - 1] = ts[adname + "FX" + str(k)][:n]
            forcing[:, len(names) + 1, k - 1] = ts[adname + "CN" + str(k)][:n]
        names = names + ["QADFX", "QADCN"]
    else:
        forcing = np.ascontiguousarray(forcing[:, :len(names), :])
    save = output_names(operation, store.act)
    with LandBatch(store, cal, names, save, device=DEVICE) as batch:
        out = batch.run(forcing)
    for r, full in enumerate(save):
        nm = full.split("_", 1)[1]
        for k in range(1, nq + 1):
            ts["%s%d_%s" % (sec, k, nm)] = np.ascontiguousarray(out[:, r, k - 1])
    errors = np.zeros(len(ERRMSGS[sec]), dtype=np.int64)
    errors[0] = int(store.dep_without_qsofg.sum())
    return errors, ERRMSGS[sec]


def pqual(io_manager, siminfo, uci, ts):
    """PQUAL.py:26-113"""
    return _qual("PERLND", siminfo, uci, ts)


def iqual(io_manager, siminfo, uci, ts):
    """IQUAL.py:24-107"""
    return _qual("IMPLND", siminfo, uci, ts)


ACTIVITIES = {
    "PERLND": {"ATEMP": atemp, "SNOW": snow, "PWATER": pwater, "SEDMNT": sedmnt, "PSTEMP": pstemp, "PWTGAS": pwtgas,
               "PQUAL": pqual},
    "IMPLND": {"ATEMP": atemp, "SNOW": snow, "IWATER": iwater, "SOLIDS": solids, "IWTGAS": iwtgas, "IQUAL": iqual},
}


def install(configuration=None, device=0):
    """Patch the reference registry in place; returns the dict of replaced callables (for uninstall)."""
    global DEVICE
    DEVICE = device
    _lib.load()  # fail loudly now if the CUDA library is missing
    if configuration is None:
        import sys

        configuration = sys.modules.get("hsp2.hsp2.configuration")
        if configuration is None:
            import importlib

            configuration = importlib.import_module("hsp2.hsp2.configuration")
    saved = {}
    for op, acts in ACTIVITIES.items():
        for name, fn in acts.items():
            saved[(op, name)] = configuration.activities[op][name]
            configuration.activities[op][name] = fn
    return saved


def uninstall(saved, configuration=None):
    if configuration is None:
        import sys

        configuration = sys.modules["hsp2.hsp2.configuration"]
    for (op, name), fn in saved.items():
        configuration.activities[op][name] = fn
