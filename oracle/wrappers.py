"""CPU-oracle activities with the reference's module contract.  TEST INFRASTRUCTURE, NOT PRODUCT.

`atemp/snow/pwater/iwater/sedmnt/pstemp/pwtgas/solids/iwtgas/pqual/iqual(io_manager, siminfo, uci, ts) -> (errors, errmsgs)` restate the
Python wrappers of the reference (ATEMP.py:16-30, SNOW.py:27-88, PWATER.py:27-99, IWATER.py:19-72,
SEDMNT.py:21-49, PSTEMP.py:23-59, PWTGAS.py:28-62) on top of the C cores in hsp2_oracle.c.  The calendar
series come from hspsquared_b200.calendar (numpy), which tests/test_calendar.py pins against series produced
by the reference's own pandas code.  `ts` is a plain dict of float64 arrays and is filled exactly like the
reference fills its typed Dict (every output name, helper arrays, NaN / zero fills) because later activities
of the same segment read them (SURVEY.md 8b "side effects to preserve").
"""
import numpy as np

from hspsquared_b200.calendar import SEASONS, Calendar

from . import oracle as O

ERRMSGS_SNOW = ("Snow simulation cannot function properly with delt> 360",)
ERRMSGS_PWATER = (
    "PWATER: Sum of irrtgt in not one", "PWATER: UZRAA exceeds UZRA array bounds",
    "PWATER: INTGB exceeds INTGRL array bounds", "PWATER: Reduced AGWO value to available",
    "PWATER: Reduced GVWS to AGWS", "PWATER: GWVS < -0.02, set to zero",
    "PWATER: Proute runoff did not converge", "PWATER: UZI highly negative", "PWATER: Reset AGWS to zero",
    "PWATER: High Water Table code not implemented")
ERRMSGS_IWATER = ("IWATER: IROUTE Newton Method did not converge",)
ERRMSGS_PSTEMP = ["SLTMP temperature less than -100C", "ULTMP temperature less than -100C",
                  "LGTMP temperature less than -100C", "SLTMP temperature greater than 100C",
                  "ULTMP temperature greater than 100C", "LGTMP temperature greater than 100C"]


def make_ui(uci):
    """utilities.py:58-79: flatten FLAGS/PARAMETERS/STATES, keeping exact Python int/float values only."""
    ui = {}
    for name in set(uci.keys()) & {"FLAGS", "PARAMETERS", "STATES"}:
        for key, value in uci[name].items():
            if type(value) in {int, float}:
                ui[key] = float(value)
    return ui


def _cal(siminfo):
    if "_calendar" not in siminfo:
        siminfo["_calendar"] = Calendar.from_siminfo(siminfo, siminfo.get("time_unit"))
    return siminfo["_calendar"]


def _metric(siminfo):
    return siminfo.get("units", 1) == 2


def _arr(ts, name):
    return np.asarray(ts[name], dtype=np.float64)


# metric units (uunits == 2): the reference converts parameters, states and parameter series on entry and the stored series
# on exit, all elementwise (PWATER.py:193-244,660-688; IWATER.py:86-131,276-283).  numpy evaluates the same IEEE operations
# in the same order as the Numba code (`x * 9./5. + 32.` is ((x * 9.) / 5.) + 32.), so the conversions live here and the C
# cores stay in English units -- except SNOW's, where `packf = PACKF + PACKI` precedes its conversion (hsp2_oracle.c).
def _f2c_in(x):
    return (x * 9. / 5.) + 32.


PWATER_OUT_METRIC = ("AGWET", "AGWI", "AGWO", "AGWS", "BASET", "CEPE", "CEPS", "GWVS", "IFWI", "IFWO", "IFWS", "IGWI", "INFIL",
                     "LZET", "LZI", "LZS", "PERC", "PERO", "PERS", "SURI", "SURO", "SURS", "TAET", "UZET", "UZI", "UZS", "SUPY", "PET")
IWATER_OUT_METRIC = ("IMPEV", "RETS", "SURI", "SURO", "SURS", "SUPY", "PET")


def atemp(io_manager, siminfo, uci, ts):
    cal = _cal(siminfo)
    ts["LAPSE"] = cal.lapse()
    ui = make_ui(uci)
    ui["k"] = siminfo["delt"] * 0.000833
    ui["steps"] = siminfo["steps"]
    return O._atemp_(ui, ts), ()


def snow(io_manager, siminfo, uci, ts):
    cal = _cal(siminfo)
    steps = siminfo["steps"]
    ts["SEASONS"] = cal.monthval(SEASONS)
    cloudfg = "CLOUD" in ts
    for name in ["AIRTMP", "CLOUD", "DTMPG", "PREC", "SOLRAD", "WINMOV"]:
        if name not in ts:
            ts[name] = np.full(steps, np.nan)
    for name in ["CCFACT", "COVIND", "MGMELT", "MWATER", "SHADE", "SNOEVP", "SNOWCF", "KMELT"]:
        if name not in ts and name in uci["PARAMETERS"]:
            ts[name] = np.full(steps, uci["PARAMETERS"][name])
    ts["HR6IND"] = cal.hour6flag(dofirst=True)
    ts["HRFG"] = cal.hoursval(np.ones(24), dofirst=True)
    siminfo["ICEFG"] = 0
    if "FLAGS" in uci and "ICEFG" in uci["FLAGS"]:
        siminfo["ICEFG"] = uci["FLAGS"]["ICEFG"]
    ui = make_ui(uci)
    ui.update(steps=steps, delt=siminfo["delt"], cloudfg=cloudfg, uunits=siminfo.get("units", 1))
    vkmfg = uci.get("FLAGS", {}).get("VKMFG", 0)
    ts["KMELT"] = cal.initm(uci, vkmfg, "MONTHLY_KMELT", uci["PARAMETERS"]["KMELT"])
    errors = O._snow_(ui, ts)
    if siminfo["delt"] > 360:
        raise KeyError("ICEFLG")  # the reference raises here (SNOW.py:85 typo); mirrored, not fixed
    return errors, ERRMSGS_SNOW


def pwater(io_manager, siminfo, uci, ts):
    cal = _cal(siminfo)
    steps = siminfo["steps"]
    for name in ("AGWLI", "IFWLI", "LGTMP", "LZLI", "PETINP", "PREC", "SURLI", "UZLI"):
        if name not in ts:
            ts[name] = np.zeros(steps)
    for name in ("AIRTMP", "PACKI", "PETINP", "PREC", "RAINF", "SNOCOV", "WYIELD"):
        if name not in ts:
            ts[name] = np.full(steps, np.nan)
    u = uci["PARAMETERS"]
    for name in ("AGWRC", "DEEPFR", "INFILT", "KVARY", "LZSN", "PETMIN", "PETMAX"):
        if name not in ts and name in u:
            ts[name] = np.full(steps, u[name])
    if "VLEFG" in u:
        flag = (u["VLEFG"] == 1) or (u["VLEFG"] == 3)
        ts["LZETP"] = cal.initm(uci, flag, "MONTHLY_LZETP", u["LZETP"])
        ts["CEPSC"] = cal.initm(uci, u["VCSFG"], "MONTHLY_CEPSC", u["CEPSC"])
        ts["INTFW"] = cal.initm(uci, u["VIFWFG"], "MONTHLY_INTFW", u["INTFW"])
        ts["IRC"] = cal.initm(uci, u["VIRCFG"], "MONTHLY_IRC", u["IRC"])
        ts["NSUR"] = cal.initm(uci, u["VNNFG"], "MONTHLY_NSUR", u["NSUR"])
        ts["UZSN"] = cal.initm(uci, u["VUZFG"], "MONTHLY_UZSN", u["UZSN"])
    else:
        for n in ("LZETP", "CEPSC", "INTFW", "IRC", "NSUR", "UZSN"):
            ts[n] = np.full(steps, u[n])
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    ts["HRFG"] = cal.hoursval(np.ones(24), dofirst=True)
    ui = make_ui(uci)
    ui.update(steps=steps, delt=siminfo["delt"])
    ui["ICEFG"] = siminfo["ICEFG"] if "ICEFG" in siminfo else 0.0
    u["CSNOFG"] = int(ui["CSNOFG"]) if "CSNOFG" in ui else 0
    if siminfo.get("units", 1) != 2:
        return O._pwater_(ui, ts), ERRMSGS_PWATER
    # ---- metric: entry conversions (PWATER.py:193-207, 236-244) on copies, the English core, exit conversions (:660-688)
    m, t2 = dict(ui), dict(ts)
    m["LSUR"] = ui["LSUR"] * 3.28
    for k in ("CEPS", "SURS", "UZS", "IFWS", "LZS", "AGWS", "GWVS"):
        m[k] = ui[k] * 0.0394
    if int(ui["ICEFG"]) and "FZG" in ui:
        m["FZG"] = ui["FZG"] * 0.0394
    for k in ("CEPSC", "UZSN", "LZSN", "WYIELD"):
        t2[k] = np.asarray(ts[k], dtype=np.float64) * 0.0394
    m["INFILT_MFAC"] = 0.0394  # INFILT = ts['INFILT'] * delt60, then * 0.0394 (PWATER.py:215, 239): the second product is the core's
    t2["KVARY"] = np.asarray(ts["KVARY"], dtype=np.float64) / 0.0394
    t2["PETMAX"] = _f2c_in(np.asarray(ts["PETMAX"], dtype=np.float64))
    t2["PETMIN"] = _f2c_in(np.asarray(ts["PETMIN"], dtype=np.float64))
    errors = O._pwater_(m, t2)
    for k in O.fields("pwater_out"):
        ts[k] = t2[k] * 25.4 if k in PWATER_OUT_METRIC else t2[k]
    return errors, ERRMSGS_PWATER


def iwater(io_manager, siminfo, uci, ts):
    cal = _cal(siminfo)
    steps = siminfo["steps"]
    for name in ("AIRTMP", "PETINP", "PREC", "RAINF", "SNOCOV", "WYIELD"):
        if name not in ts:
            ts[name] = np.full(steps, np.nan)
    if "SURLI" not in ts:
        ts["SURLI"] = np.zeros(steps)
    u = uci["PARAMETERS"]
    for name in ["PETMAX", "PETMIN"]:
        if name not in ts:
            ts[name] = np.full(steps, u[name], dtype=np.float64)
    if "VRSFG" in u:
        ts["RETSC"] = cal.initm(uci, u["VRSFG"], "MONTHLY_RETSC", u["RETSC"])
        ts["NSUR"] = cal.initm(uci, u["VNNFG"], "MONTHLY_NSUR", u["NSUR"])
    else:
        ts["RETSC"] = np.full(steps, u["RETSC"])
        ts["NSUR"] = np.full(steps, u["NSUR"])
    ts["HR1FG"] = cal.hourflag(1, dofirst=True)
    ts["HRFG"] = cal.hoursval(np.ones(24), dofirst=True)
    ui = make_ui(uci)
    ui.update(steps=steps, delt=siminfo["delt"])
    if siminfo.get("units", 1) != 2:
        return O._iwater_(ui, ts), ERRMSGS_IWATER
    # ---- metric: IWATER.py:86-87, 112-116, 131-133 on entry, :276-283 on exit
    m, t2 = dict(ui), dict(ts)
    m["LSUR"] = ui["LSUR"] * 3.28
    m["RETS"], m["SURS"] = ui["RETS"] * 0.0394, ui["SURS"] * 0.0394
    t2["RETSC"] = np.asarray(ts["RETSC"], dtype=np.float64) * 0.0394
    t2["WYIELD"] = np.asarray(ts["WYIELD"], dtype=np.float64) * 0.0394
    t2["PETMAX"] = _f2c_in(np.asarray(ts["PETMAX"], dtype=np.float64))
    t2["PETMIN"] = _f2c_in(np.asarray(ts["PETMIN"], dtype=np.float64))
    errors = O._iwater_(m, t2)
    for k in O.fields("iwater_out"):
        ts[k] = t2[k] * 25.4 if k in IWATER_OUT_METRIC else t2[k]
    return errors, ERRMSGS_IWATER


def sedmnt(io_manager, siminfo, uci, ts):
    cal = _cal(siminfo)
    n = siminfo["steps"]
    ui = make_ui(uci)
    ui.update(simlen=n, delt=siminfo["delt"], uunits=siminfo.get("units", 1))
    u = uci["PARAMETERS"]
    ts["COVERI"] = cal.initm(uci, u["CRVFG"], "MONTHLY_COVER", u["COVER"]) if "CRVFG" in u else np.full(n, u["COVER"])
    ts["NVSI"] = cal.initm(uci, u["VSIVFG"], "MONTHLY_NVSI", u["NVSI"]) if "VSIVFG" in u else np.full(n, u["NVSI"])
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    if not _metric(siminfo):
        return O._sedmnt_(ui, ts), []
    # ---- metric (SEDMNT.py:113-117): SURO / SURS back to inches on local copies; NVSI, DETS and the stored series in the core
    for k in ("PREC", "SLSED", "SNOCOV", "SURO", "SURS"):  # SEDMNT.py:94-96: the zero fills stay in ts
        if k not in ts:
            ts[k] = np.zeros(n)
    t2 = dict(ts)
    t2["SURO"], t2["SURS"] = _arr(ts, "SURO") * 0.0394, _arr(ts, "SURS") * 0.0394
    errors = O._sedmnt_(ui, t2)
    for k in O.fields("sedmnt_out"):
        ts[k] = t2[k]
    return errors, []


def solids(io_manager, siminfo, uci, ts):
    """SOLIDS.py:16-49"""
    cal = _cal(siminfo)
    n = siminfo["steps"]
    for nm in ("SURO", "SURS", "PREC", "SLSLD"):
        if nm not in ts:
            ts[nm] = np.zeros(n)
    u = uci["PARAMETERS"]
    ts["ACCSDP"] = cal.initm(uci, u["VASDFG"], "MONTHLY_ACCSDP", u["ACCSDP"]) if "VASDFG" in u else np.full(n, u["ACCSDP"])
    ts["REMSDP"] = cal.initm(uci, u["VRSDFG"], "MONTHLY_REMSDP", u["REMSDP"]) if "VRSDFG" in u else np.full(n, u["REMSDP"])
    ui = make_ui(uci)
    ui.update(simlen=n, delt60=siminfo["delt"] / 60)
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    if not _metric(siminfo):
        return O._solids_(ui, ts), []
    # ---- metric: SOLIDS.py:84-88 on entry (local copies), :143-145 on exit
    m, t2 = dict(ui), dict(ts)
    t2["ACCSDP"] = _arr(ts, "ACCSDP") * 1.10231 / 2.471
    t2["SURO"], t2["SURS"] = _arr(ts, "SURO") * 0.0394, _arr(ts, "SURS") * 0.0394
    m["SLDS"] = ui["SLDS"] * 1.10231 / 2.471
    errors = O._solids_(m, t2)
    for k in O.fields("solids_out"):
        ts[k] = t2[k] / 1.10231 * 2.471
    return errors, []


def iwtgas(io_manager, siminfo, uci, ts):
    """IWTGAS.py:33-63"""
    cal = _cal(siminfo)
    n = siminfo["steps"]
    ui = make_ui(uci)
    ui.update(simlen=n, uunits=siminfo.get("units", 1))
    u = uci["PARAMETERS"]
    for nm in ("AWTF", "BWTF"):
        ts[nm] = cal.initm(uci, u["WTFVFG"], "MONTHLY_" + nm, u[nm]) if "WTFVFG" in u else np.full(n, u[nm])
    for nm in ("AIRTMP", "WYIELD", "SURO", "SURLI"):
        if nm not in ts:
            ts[nm] = np.zeros(n)
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    if not _metric(siminfo):
        return O._iwtgas_(ui, ts), []
    # ---- metric: IWTGAS.py:80-82 (ELEV; SOTMP is overwritten before it is read), :111-114 on local copies; the stored series in the core
    m, t2 = dict(ui), dict(ts)
    m["ELEV"] = ui["ELEV"] * 3.281
    t2["AWTF"] = _f2c_in(_arr(ts, "AWTF"))
    t2["WYIELD"], t2["SURO"] = _arr(ts, "WYIELD") * 0.0394, _arr(ts, "SURO") * 0.0394
    errors = O._iwtgas_(m, t2)
    for k in list(O.fields("iwtgas_out")) + ["SLITMP", "SLIDOX", "SLICO2"]:
        ts[k] = t2[k]
    return errors, []


def pstemp(io_manager, siminfo, uci, ts):
    cal = _cal(siminfo)
    n = siminfo["steps"]
    ui = make_ui(uci)
    ui.update(simlen=n, delt=siminfo["delt"], uunits=siminfo.get("units", 1))  # metric: inside the core (hsp2_oracle.c)
    u = uci["PARAMETERS"]
    for flag, names in (("SLTVFG", ("ASLT", "BSLT")), ("ULTVFG", ("ULTP1", "ULTP2")), ("LGTVFG", ("LGTP1", "LGTP2"))):
        for nm in names:
            ts[nm] = cal.initm(uci, u[flag], "MONTHLY_" + nm, u[nm]) if flag in u else np.full(n, u[nm])
    ts["HRFG"] = cal.hoursval(np.ones(24), dofirst=True)
    return O._pstemp_(ui, ts), ERRMSGS_PSTEMP


def pwtgas(io_manager, siminfo, uci, ts):
    cal = _cal(siminfo)
    n = siminfo["steps"]
    ui = make_ui(uci)
    ui.update(simlen=n, uunits=siminfo.get("units", 1))
    u = uci["PARAMETERS"]
    for flag, nm in (("IDVFG", "IDOXP"), ("ICVFG", "ICO2P"), ("GDVFG", "ADOXP"), ("GCVFG", "ACO2P")):
        ts[nm] = cal.initm(uci, u[flag], "MONTHLY_" + nm, u[nm]) if flag in u else np.full(n, u[nm])
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    if not _metric(siminfo):
        return O._pwtgas_(ui, ts), []
    # ---- metric: PWTGAS.py:90-94 (ELEV; SOTMP / IOTMP / AOTMP are overwritten before they are read), :167-174 on local copies
    # -- the -1e30 fills of absent soil temperatures are converted like any other value --; the stored series in the core
    m, t2 = dict(ui), dict(ts)
    m["ELEV"] = ui["ELEV"] * 3.281
    fills = {}
    for k in ("WYIELD", "SURO", "IFWO", "AGWO", "SURLI", "IFWLI", "AGWLI"):  # PWTGAS.py:106-108
        if k not in ts:
            fills[k] = np.zeros(n)
    for k in ("SLTMP", "ULTMP", "LGTMP", "SLITMP", "SLIDOX", "SLICO2", "ILITMP", "ILIDOX", "ILICO2", "ALITMP", "ALIDOX", "ALICO2"):
        if k not in ts:
            fills[k] = np.full(n, -1.0e30)
    ts.update(fills)
    t2.update(fills)
    for k in ("SLTMP", "ULTMP", "LGTMP"):
        t2[k] = _f2c_in(_arr(ts, k))
    for k in ("WYIELD", "SURO", "IFWO", "AGWO"):
        t2[k] = _arr(ts, k) * 0.0394
    errors = O._pwtgas_(m, t2)
    for k in O.fields("pwtgas_out"):
        ts[k] = t2[k]
    return errors, []


ERRMSGS_PQUAL = ("PQUAL: A constituent must be associated with overland flow in order to receive atmospheric deposition inputs", "")
ERRMSGS_IQUAL = ("IQUAL: A constituent must be associated with overland flow in order to receive atmospheric deposition inputs", "")


def _initmdiv(cal, uci, flag, monthly1, monthly2, default1, default2):
    """utilities.py:214-225: the ACQOP / SQOLIM special case (monthly tables divided month by month)."""
    if flag and monthly1 and monthly2 in uci:
        m1, m2 = list(uci[monthly1].values()), list(uci[monthly2].values())
        return cal.dayval(np.array([m1[m] / m2[m] for m in range(12)]))
    return np.full(cal.steps, default1 / default2)


def _qual_common(pre, siminfo, uci, ts, monthly_names, adname):
    """The part of pqual() / iqual() that is the same (PQUAL.py:32-108, IQUAL.py:29-99)."""
    cal = _cal(siminfo)
    n = siminfo["steps"]
    nquals = 1
    if "PARAMETERS" in uci and "NQUAL" in uci["PARAMETERS"]:
        nquals = int(uci["PARAMETERS"]["NQUAL"])
    for k in range(nquals):
        uci[pre + str(k + 1) + "_FLAGS"]["QUALID"]  # KeyError like the reference when a constituent's tables are missing
    ui = make_ui(uci)
    ui.update(simlen=n, delt60=siminfo["delt"] / 60, nquals=nquals, errlen=2)
    u = uci.get("FLAGS")
    for k in range(1, nquals + 1):
        K = str(k)
        f, p = uci[pre + K + "_FLAGS"], uci[pre + K + "_PARAMETERS"]
        for key, flagname in monthly_names:
            tab = pre + K + "_MONTHLY/" + ("IOQC" if key == "IOQCP" else "AOQC" if key == "AOQCP" else key)
            ts[key + K] = cal.initm(uci, f[flagname], tab, p["IOQC" if key == "IOQCP" else "AOQC" if key == "AOQCP" else key])
        ts["REMQOP" + K] = _initmdiv(cal, uci, f["VQOFG"], pre + K + "_MONTHLY/ACQOP", pre + K + "_MONTHLY/SQOLIM",
                                     p["ACQOP"], p["SQOLIM"])
        adf = adc = 0
        ts[adname + "FX" + K] = np.zeros(n)
        ts[adname + "CN" + K] = np.zeros(n)
        if u is not None:
            adf = u[adname + "FG" + str(2 * k - 1)]
            if adf > 0:
                ts[adname + "FX" + K] = cal.initm(uci, adf, pre + K + "_MONTHLY/" + adname + "FX", 0.0)
            elif adf == -1:
                ts[adname + "FX" + K] = ts[adname + "FX" + K + " 1"]
            adc = u[adname + "FG" + str(2 * k)]
            if adc > 0:
                ts[adname + "CN" + K] = cal.initm(uci, adc, pre + K + "_MONTHLY/" + adname + "CN", 0.0)
            elif adc == -1:
                ts[adname + "CN" + K] = ts[adname + "CN" + K + " 1"]
        ui[adname.lower() + "fgf" + K] = adf
        ui[adname.lower() + "fgc" + K] = adc
    ts["DAYFG"] = cal.hourflag(0, dofirst=True)
    return ui, nquals, n


def pqual(io_manager, siminfo, uci, ts):
    """PQUAL.py:26-113 -- no unit handling in the reference: a metric run reads the flows in the units they were stored in"""
    ui, nquals, n = _qual_common(
        "PQUAL", siminfo, uci, ts,
        (("POTFW", "VPFWFG"), ("POTFS", "VPFSFG"), ("ACQOP", "VQOFG"), ("SQOLIM", "VQOFG"), ("IOQCP", "VIQCFG"), ("AOQCP", "VAQCFG")),
        "PQAD")
    for k in range(1, nquals + 1):
        K = str(k)
        f, p = uci["PQUAL" + K + "_FLAGS"], uci["PQUAL" + K + "_PARAMETERS"]
        for nm in ("QSDFG", "QSOFG", "QIFWFG", "QAGWFG", "VIQCFG", "VAQCFG"):
            ui[nm + K] = f[nm]
        ui["SQO" + K], ui["WSQOP" + K] = p["SQO"], p["WSQOP"]
    for nm in ("SLIQSP", "ILIQC", "ALIQC", "SURO", "IFWO", "AGWO", "PERO", "WSSD", "SCRSD"):  # PQUAL.py:98-106
        if nm not in ts:
            ts[nm] = np.zeros(n)
    return O._pqual_(ui, ts), ERRMSGS_PQUAL


def iqual(io_manager, siminfo, uci, ts):
    """IQUAL.py:24-107 -- no unit handling in the reference (see pqual)"""
    ui, nquals, n = _qual_common("IQUAL", siminfo, uci, ts, (("POTFW", "VPFWFG"), ("ACQOP", "VQOFG"), ("SQOLIM", "VQOFG")),
                                 "IQAD")
    for k in range(1, nquals + 1):
        K = str(k)
        f, p = uci["IQUAL" + K + "_FLAGS"], uci["IQUAL" + K + "_PARAMETERS"]
        for nm in ("QSDFG", "QSOFG", "VQOFG"):
            ui[nm + K] = f[nm]
        ui["sqo" + K], ui["wsqop" + K] = p["SQO"], p["WSQOP"]
    for nm in ("SLIQSX", "SLIQO", "SLIQSP"):  # IQUAL.py:96-98
        if nm not in ts:
            ts[nm] = np.full(n, -1.0e30)
    return O._iqual_(ui, ts), ERRMSGS_IQUAL
