"""Parity cases shared by oracle/gen_golden.py (reference side, build container only) and tests/ (oracle and
CUDA side).  TEST INFRASTRUCTURE.

A case = (operation, per-activity UCI tables, time base, forcing recipe).  `run_segment` walks the activities
of one segment exactly like the reference's dispatcher does (main.py:123-146, 249-250): registry order,
skip when the ACTIVITY flag is 0, CSNOFG -> SEDMNT/PWTGAS and AIRTFG -> PSTEMP injections.
"""
import copy
import os

import numpy as np

from hspsquared_b200 import cbp, test10
from hspsquared_b200.test10 import monthly

ORDER = {
    "PERLND": ("ATEMP", "SNOW", "PWATER", "SEDMNT", "PSTEMP", "PWTGAS", "PQUAL"),  # configuration.py:48-50
    "IMPLND": ("ATEMP", "SNOW", "IWATER", "SOLIDS", "IWTGAS", "IQUAL"),  # configuration.py:51-52
}

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hspsquared_b200", "data")


def run_segment(operation, tables, ts, siminfo, activities, io_manager=None):
    """-> {activity: errors}; outputs land in ts.  `activities`: {name: callable(io, siminfo, uci, ts)}."""
    flags = tables["ACTIVITY"]
    errs = {}
    for act in ORDER[operation]:
        if act not in activities or not flags.get(act, 0):
            continue
        ui = tables[act]
        if operation == "PERLND" and act in ("SEDMNT", "PWTGAS"):
            ui["PARAMETERS"]["CSNOFG"] = tables["PWATER"]["PARAMETERS"]["CSNOFG"]
        if operation == "PERLND" and act == "PSTEMP":
            ui["PARAMETERS"]["AIRTFG"] = flags["ATEMP"]
        e, _msgs = activities[act](io_manager, siminfo, ui, ts)
        errs[act] = np.asarray(e, dtype=np.int64)
    return errs


# ---------------------------------------------------------------------------------------------------
# forcing recipes (deterministic numpy, from the committed 1976 test10 met decoded by gen_golden.py)
# ---------------------------------------------------------------------------------------------------
_met_cache = {}


def load_met():
    if "met" not in _met_cache:
        z = np.load(os.path.join(DATA, "test10_met_1976.npz"))
        _met_cache["met"] = {k: z[k] for k in z.files}
    return _met_cache["met"]


FLOWS = ("PREC", "PETINP", "WINMOV", "SOLRAD")  # amounts per interval: split / summed on resampling


def forcings(case):
    """Hourly 1976 series -> the case's window and step.  Returns dict name -> float64[steps]."""
    met = load_met()
    pre = "P_" if case["met"] == "perlnd" else "I_"
    names = ["PREC", "AIRTMP", "PETINP", "WINMOV", "SOLRAD", "DTMPG", "CLOUD"]
    hourly = {n: met[pre + n].copy() for n in names}
    if case["met"] == "flowtest":
        hourly = {"PREC": met["F_PREC"].copy(), "PETINP": met["F_PETINP"].copy()}
    h0 = case.get("hour0", 0)  # first hour of 1976 used
    delt = case["delt"]
    steps = case["steps"]
    out = {}
    for n, a in hourly.items():
        if delt == 60:
            v = a[h0:h0 + steps]
        elif delt < 60:
            k = 60 // delt
            nh = (steps + k - 1) // k
            v = np.repeat(a[h0:h0 + nh], k)[:steps]
            if n in FLOWS:
                v = v / k
        else:
            k = delt // 60
            blk = a[h0:h0 + steps * k].reshape(steps, k)
            v = blk.sum(axis=1) if n in FLOWS else blk[:, 0].copy()
        out[n] = np.ascontiguousarray(v, dtype=np.float64)
    for n in case.get("drop", ()):
        out.pop(n, None)
    rng = np.random.default_rng(case.get("seed", 7))
    if case.get("rain_mult"):
        out["PREC"] = out["PREC"] * case["rain_mult"]
    if "gatmp" in case:  # ATEMP active: gauge temperature in, AIRTMP computed
        out["GATMP"] = out.pop("AIRTMP") + case["gatmp"]
    for n in case.get("lateral", ()):  # sparse positive lateral inflow series
        a = rng.random(steps)
        out[n] = np.where(a > 0.97, 0.02 * rng.random(steps), 0.0)
    for n, (lo, hi) in case.get("lateral_conc", {}).items():
        out[n] = lo + (hi - lo) * rng.random(steps)
    if case.get("qual_inputs"):  # PQUAL alone: every flow and sediment series it reads is an external series
        wet = out["PREC"] > 0.0
        out["SURO"] = np.where(wet, 0.6 * out["PREC"], 0.0) * rng.random(steps)
        out["IFWO"] = 0.002 * rng.random(steps) * (rng.random(steps) > 0.3)
        out["AGWO"] = 0.001 + 0.001 * rng.random(steps)
        out["PERO"] = out["SURO"] + out["IFWO"] + out["AGWO"]
        out["WSSD"] = np.where(out["SURO"] > 0.0, 0.05 * rng.random(steps), 0.0)
        out["SCRSD"] = np.where(out["SURO"] > 0.01, 0.01 * rng.random(steps), 0.0)
    for key, scale in case.get("dep_series", {}).items():  # atmospheric deposition given as an input series (flag -1)
        out[key] = scale * rng.random(steps) * (rng.random(steps) > 0.5)
    if case.get("runoff_inputs"):  # SURO / SURS given as external series (the sections that make them are off)
        out["SURO"] = np.where(out["PREC"] > 0.0, 0.8 * out["PREC"], 0.0) * rng.random(steps)
        out["SURS"] = 0.05 * rng.random(steps)
    return out


def siminfo_for(case):
    start = np.datetime64("%04d-01-01T00:00" % case.get("year", 1976)) + np.timedelta64(case.get("hour0", 0), "h")
    stop = start + np.timedelta64(case["steps"] * case["delt"], "m")
    return {"start": start, "stop": stop, "delt": case["delt"], "units": case.get("units", 1), "steps": case["steps"]}


# ---------------------------------------------------------------------------------------------------
# table variants
# ---------------------------------------------------------------------------------------------------
def _cbp_tables():
    return cbp.perlnd()


def _set(tables, act, table, **kv):
    tables[act].setdefault(table, {}).update(kv)
    return tables


def build_cases():
    """name -> case dict (tables + recipe).  Everything is deterministic."""
    C = {}
    Y = 8784  # hourly steps in 1976

    def add(name, op, tables, met, steps=Y, delt=60, **kw):
        C[name] = dict(name=name, operation=op, tables=tables, met=met, steps=steps, delt=delt, **kw)

    # -- the two test10 land segments as shipped (CLOUD present) ---------------------------------
    add("p001_test10", "PERLND", test10.perlnd(), "perlnd")
    add("i001_test10", "IMPLND", test10.implnd(), "implnd")

    # -- BASELINE config 2 base: SNOW+PWATER only, 6 forcings (no CLOUD) -------------------------
    t = test10.perlnd()
    t["ACTIVITY"].update({"PSTEMP": 0, "PWTGAS": 0})
    add("p001_snow_pwater", "PERLND", t, "perlnd", drop=("CLOUD",))

    # -- full chain, CBP-style (ATEMP on, RTOPFG=1, SDOPFG=1, TSOPFG=1, monthly everything) -------
    add("cbp_fullchain", "PERLND", _cbp_tables(), "perlnd", drop=("CLOUD",), gatmp=2.0)

    # -- full chain variants: UZFG=1 + Newton routing + method-2 washoff + TSOPFG=0 + VSIVFG=2 ----
    t = _cbp_tables()
    t["PWATER"]["PARAMETERS"].update({"RTOPFG": 0, "UZFG": 1, "VLEFG": 2, "KVARY": 0.0, "BASETP": 0.15})
    t["SEDMNT"]["PARAMETERS"].update({"SDOPFG": 0, "VSIVFG": 2, "KGER": 0.3, "JGER": 1.5})
    t["PSTEMP"]["PARAMETERS"].update({"TSOPFG": 0, "SLTVFG": 0, "ULTVFG": 0, "LGTVFG": 0, "ASLT": 14.5,
                                      "BSLT": 0.365, "ULTP1": 0.1, "ULTP2": 4.0, "LGTP1": 0.05, "LGTP2": 6.0})
    t["PWTGAS"]["PARAMETERS"].update({"IDVFG": 0, "GDVFG": 0, "ICVFG": 1, "GCVFG": 1, "IDOXP": 6.0, "ADOXP": 5.0})
    t["PWTGAS"]["MONTHLY_ICO2P"] = monthly([.05, .05, .06, .07, .08, .1, .12, .12, .1, .08, .06, .05])
    t["PWTGAS"]["MONTHLY_ACO2P"] = monthly([.04, .04, .05, .05, .06, .07, .08, .08, .07, .06, .05, .04])
    add("chain_uzfg_newton", "PERLND", t, "perlnd", drop=("CLOUD",), gatmp=-1.5, rain_mult=3.0)

    # -- TSOPFG=2, VSIVFG=1, ICEFG=0, IFRDFG=1, VLEFG=3 ------------------------------------------
    t = _cbp_tables()
    t["ACTIVITY"]["ATEMP"] = 0
    t["SNOW"]["FLAGS"]["ICEFG"] = 0
    t["PWATER"]["PARAMETERS"].update({"IFRDFG": 1, "VLEFG": 3, "VIFWFG": 1, "VIRCFG": 1})
    t["PWATER"]["MONTHLY_INTFW"] = monthly([2., 2., 2.2, 2.5, 2.5, 2.2, 2., 1.8, 1.8, 2., 2., 2.])
    t["PWATER"]["MONTHLY_IRC"] = monthly([.45, .45, .5, .55, .6, .6, .55, .5, .5, .45, .45, .45])
    t["SEDMNT"]["PARAMETERS"].update({"VSIVFG": 1})
    t["PSTEMP"]["PARAMETERS"].update({"TSOPFG": 2, "SLTVFG": 0, "ULTVFG": 0, "LGTVFG": 0, "ASLT": 14.5,
                                      "BSLT": 0.365, "ULTP1": 0.1, "ULTP2": 4.0, "LGTP1": 0.05, "LGTP2": 6.0})
    add("chain_tsop2_ifrd", "PERLND", t, "perlnd")

    # -- degree-day snow (SNOPFG=1) with monthly KMELT ------------------------------------------
    t = test10.perlnd()
    t["ACTIVITY"].update({"PSTEMP": 0, "PWTGAS": 0})
    t["SNOW"]["FLAGS"].update({"SNOPFG": 1, "VKMFG": 1, "ICEFG": 0})
    t["SNOW"]["PARAMETERS"]["KMELT"] = 0.05
    t["SNOW"]["MONTHLY_KMELT"] = monthly([.03, .03, .05, .08, .1, .1, .1, .1, .08, .06, .04, .03])
    add("p_degday", "PERLND", t, "perlnd", drop=("CLOUD",))

    # -- no SNOW at all: CSNOFG=0, IFFCFG=2 quirk, SEDMNT sees NaN RAINF -------------
    t = _cbp_tables()
    t["ACTIVITY"].update({"ATEMP": 0, "SNOW": 0})
    t["PWATER"]["PARAMETERS"].update({"CSNOFG": 0, "IFFCFG": 2})
    add("p_nosnow_iffc2", "PERLND", t, "perlnd")

    # -- sub-hourly and super-hourly steps ------------------------------------------------------
    t = test10.perlnd()
    add("p001_delt30", "PERLND", t, "perlnd", steps=2 * 24 * 120, delt=30)
    add("cbp_delt15", "PERLND", _cbp_tables(), "perlnd", steps=4 * 24 * 75, delt=15, gatmp=1.0, hour0=24 * 40)
    t = test10.perlnd()
    add("p001_delt120", "PERLND", t, "perlnd", steps=12 * 200, delt=120)

    # -- run starting mid-year, not on a day boundary (calendar dofirst quirk) -------------------
    add("p001_midyear", "PERLND", test10.perlnd(), "perlnd", steps=24 * 150 + 7, hour0=24 * 74 + 6)

    # -- stress: 8x rain, tiny UZSN (table-bounds errors), thin soil ------------------------------
    t = test10.perlnd()
    t["ACTIVITY"].update({"PSTEMP": 0, "PWTGAS": 0})
    t["PWATER"]["PARAMETERS"].update({"VUZFG": 0, "UZSN": 0.02, "INFILT": 0.01, "LZSN": 2.0, "AGWRC": 0.999,
                                      "KVARY": 3.0, "NSUR": 0.4})
    t["PWATER"]["STATES"].update({"UZS": 0.5, "AGWS": 0.0, "GWVS": 0.0})
    add("p_stress_wet", "PERLND", t, "perlnd", rain_mult=8.0)

    # -- lateral inflows + lateral concentrations -------------------------------------------------
    t = _cbp_tables()
    t["PWTGAS"]["PARAMETERS"].update({"SLIFAC": 0.5, "ILIFAC": 0.3, "ALIFAC": 0.2})
    add("cbp_lateral", "PERLND", t, "perlnd", gatmp=0.5, steps=24 * 200,
        lateral=("SURLI", "UZLI", "IFWLI", "LZLI", "AGWLI", "SLSED"),
        lateral_conc={"SLITMP": (35., 70.), "SLIDOX": (4., 12.), "SLICO2": (0., .2), "ILITMP": (35., 70.),
                      "ILIDOX": (4., 12.), "ILICO2": (0., .2), "ALITMP": (40., 60.), "ALIDOX": (3., 9.),
                      "ALICO2": (0., .2)})

    # -- IMPLND variants --------------------------------------------------------------------------
    t = test10.implnd()
    t["ACTIVITY"]["SNOW"] = 0
    t["IWATER"]["PARAMETERS"].update({"CSNOFG": 0, "RTOPFG": 1, "VRSFG": 1, "VNNFG": 1, "RTLIFG": 0})
    t["IWATER"]["MONTHLY_RETSC"] = monthly([.01, .01, .02, .03, .05, .06, .06, .05, .04, .02, .01, .01])
    t["IWATER"]["MONTHLY_NSUR"] = monthly([.01, .01, .012, .015, .02, .02, .02, .02, .015, .012, .01, .01])
    add("i_rtop_monthly", "IMPLND", t, "implnd", lateral=("SURLI",))
    add("i001_delt15", "IMPLND", test10.implnd(), "implnd", steps=4 * 24 * 60, delt=15, hour0=24 * 20)
    t = test10.implnd()
    t["ACTIVITY"]["ATEMP"] = 1
    t["ATEMP"] = {"PARAMETERS": {"ELDAT": -300.0, "AIRTMP": 30.0}}
    add("i_atemp_snow", "IMPLND", t, "implnd", gatmp=0.0, drop=("CLOUD",), rain_mult=2.0)

    # -- IMPLND SOLIDS + IWTGAS: test10's chain with the IWT-INIT table the shipped UCI lacks ---------
    t = test10.implnd()
    t["ACTIVITY"]["IWTGAS"] = 1
    t["IWTGAS"]["STATES"] = {"SOTMP": 60.0, "SODOX": 0.0, "SOCO2": 0.0}  # ParseTable.csv rows 1360-1362 defaults
    add("i001_iwtgas", "IMPLND", t, "implnd")
    # washoff method 1, monthly accumulation / removal / regression coefficients, ATEMP, lateral inflow with
    # temperature and DO but no CO2 concentration (the nested mixing of IWTGAS.py:151-158)
    t = test10.implnd()
    t["ACTIVITY"].update({"ATEMP": 1, "IWTGAS": 1})
    t["ATEMP"] = {"PARAMETERS": {"ELDAT": 120.0, "AIRTMP": 30.0}}
    t["SOLIDS"]["PARAMETERS"].update({"SDOPFG": 1, "VASDFG": 1, "VRSDFG": 1, "KEIM": 0.3, "JEIM": 1.4})
    t["SOLIDS"]["MONTHLY_ACCSDP"] = monthly([.004, .004, .006, .01, .012, .014, .014, .012, .01, .008, .006, .004])
    t["SOLIDS"]["MONTHLY_REMSDP"] = monthly([.03, .03, .04, .05, .06, .07, .07, .06, .05, .04, .03, .03])
    t["IWTGAS"]["PARAMETERS"].update({"WTFVFG": 1, "SLIFAC": 0.4, "ELEV": 1250.0})
    t["IWTGAS"]["STATES"] = {"SOTMP": 60.0, "SODOX": 0.0, "SOCO2": 0.0}
    t["IWTGAS"]["MONTHLY_AWTF"] = monthly([34., 35., 38., 44., 52., 60., 66., 65., 58., 48., 40., 35.])
    t["IWTGAS"]["MONTHLY_BWTF"] = monthly([.3, .3, .35, .4, .45, .5, .5, .5, .45, .4, .35, .3])
    add("i_solids_m1_wtf_monthly", "IMPLND", t, "implnd", gatmp=1.0, drop=("CLOUD",), rain_mult=2.5, lateral=("SURLI",),
        lateral_conc={"SLITMP": (35., 70.), "SLIDOX": (4., 12.)})
    # SOLIDS / IWTGAS alone on given runoff series (no SNOW, no IWATER in the batch): zero / -1e30 fills
    t = test10.implnd()
    t["ACTIVITY"].update({"SNOW": 0, "IWATER": 0, "IWTGAS": 1})
    t["IWTGAS"]["STATES"] = {"SOTMP": 60.0, "SODOX": 0.0, "SOCO2": 0.0}
    add("i_solids_iwtgas_alone", "IMPLND", t, "implnd", steps=24 * 120, runoff_inputs=True)

    # -- IQUAL: test10's I001 as shipped (one constituent, COD: QSDFG=1, QSOFG=1; test10.uci:274-290) --------------
    t = test10.implnd()
    t["ACTIVITY"]["IQUAL"] = 1
    t["IQUAL"] = test10.iqual_tables()
    add("i001_iqual", "IMPLND", t, "implnd")
    # three constituents: QSOFG=2 with monthly accumulation / limit and atmospheric deposition (monthly flux and
    # concentration); the constant-concentration special case QSOFG=-1; solids-associated with monthly potency, lateral potency
    t = test10.implnd()
    t["ACTIVITY"]["IQUAL"] = 1
    t["IQUAL"] = {
        "PARAMETERS": {"NQUAL": 3, "SDLFAC": 0.0, "SLIFAC": 0.35},
        "FLAGS": {"IQADFG%d" % k: 0 for k in range(1, 21)},
        "IQUAL1_FLAGS": {"QUALID": "NO3", "QTYID": "LB", "QSDFG": 0, "VPFWFG": 0, "QSOFG": 2, "VQOFG": 1},
        "IQUAL1_PARAMETERS": {"SQO": 0.8, "POTFW": 0.0, "ACQOP": 0.03, "SQOLIM": 1.5, "WSQOP": 0.9},
        "IQUAL1_MONTHLY/ACQOP": monthly([.02, .02, .025, .03, .035, .04, .04, .04, .035, .03, .025, .02]),
        "IQUAL1_MONTHLY/SQOLIM": monthly([1., 1., 1.2, 1.4, 1.6, 1.8, 1.8, 1.8, 1.6, 1.4, 1.2, 1.]),
        "IQUAL1_MONTHLY/IQADFX": monthly([.001, .001, .002, .002, .003, .003, .003, .003, .002, .002, .001, .001]),
        "IQUAL1_MONTHLY/IQADCN": monthly([.01, .01, .012, .014, .016, .018, .018, .016, .014, .012, .01, .01]),
        "IQUAL2_FLAGS": {"QUALID": "CL", "QTYID": "LB", "QSDFG": 0, "VPFWFG": 0, "QSOFG": -1, "VQOFG": 0},
        "IQUAL2_PARAMETERS": {"SQO": 0.0, "POTFW": 0.0, "ACQOP": 12.0, "SQOLIM": 1.0, "WSQOP": 1.64},
        "IQUAL3_FLAGS": {"QUALID": "PB", "QTYID": "LB", "QSDFG": 1, "VPFWFG": 1, "QSOFG": 0, "VQOFG": 0},
        "IQUAL3_PARAMETERS": {"SQO": 0.4, "POTFW": 0.2, "ACQOP": 0.0, "SQOLIM": 0.000001, "WSQOP": 1.64},
        "IQUAL3_MONTHLY/POTFW": monthly([.1, .1, .12, .15, .2, .25, .25, .25, .2, .15, .12, .1]),
    }
    t["IQUAL"]["FLAGS"].update({"IQADFG1": 1, "IQADFG2": 2})
    add("i_iqual_variants", "IMPLND", t, "implnd", rain_mult=2.0, lateral_conc={"SLIQSP": (-0.2, 0.6)})

    # -- PQUAL on the CBP full chain: sediment-associated with monthly potency; overland-flow-associated with monthly
    #    accumulation / limit + interflow and groundwater concentrations (monthly, mg/l flavour VAQCFG=3); QSOFG=2 with
    #    atmospheric deposition, constant concentrations, lateral concentrations ------------------------------------------
    t = _cbp_tables()
    t["ACTIVITY"]["PQUAL"] = 1
    t["PQUAL"] = test10.pqual_tables()
    add("cbp_pqual", "PERLND", t, "perlnd", drop=("CLOUD",), gatmp=2.0, rain_mult=2.0,
        lateral_conc={"SLIQSP": (-0.3, 1.0), "ILIQC": (-5.0, 20.0), "ALIQC": (-5.0, 20.0)})
    # PQUAL alone on external flow / sediment series; constituent 2 takes deposition without QSOFG (errors[0] = 1)
    t = {"ACTIVITY": {"PQUAL": 1}, "PQUAL": test10.pqual_tables()}
    t["PQUAL"]["PQUAL2_FLAGS"].update({"QSOFG": 0})
    t["PQUAL"]["FLAGS"].update({"PQADFG3": 1})
    t["PQUAL"]["PQUAL2_MONTHLY/PQADFX"] = monthly([.001] * 12)
    add("p_pqual_alone", "PERLND", t, "perlnd", steps=24 * 150, qual_inputs=True)
    # ... and atmospheric deposition as INPUT SERIES (PQADFG = -1, PQUAL.py:94-103): constituent 3 takes its dry flux from a
    # series (its wet concentration stays monthly), constituent 2 its wet concentration
    t = {"ACTIVITY": {"PQUAL": 1}, "PQUAL": test10.pqual_tables()}
    t["PQUAL"]["FLAGS"].update({"PQADFG5": -1, "PQADFG6": 1, "PQADFG4": -1})
    add("p_pqual_depseries", "PERLND", t, "perlnd", steps=24 * 120, qual_inputs=True,
        dep_series={"PQADFX3 1": 0.004, "PQADCN2 1": 0.02})

    # -- the reference's own unit-test cases (tests/ipwater/test_ipwater.py:43-84) ----------------
    fp = {"ACTIVITY": {"PWATER": 1}, "PWATER": copy.deepcopy(FLOWTEST_PERLND)}
    add("flowtest_pwater", "PERLND", fp, "flowtest", steps=8760, year=2018)
    fi = {"ACTIVITY": {"IWATER": 1}, "IWATER": copy.deepcopy(FLOWTEST_IMPLND)}
    add("flowtest_iwater", "IMPLND", fi, "flowtest", steps=8760, year=2018)

    # -- metric units (siminfo['units'] == 2): entry / exit conversions of SNOW, PWATER, IWATER (SNOW.py:114-160,569-585;
    #    PWATER.py:193-244,660-688; IWATER.py:86-131,276-283).  The numbers are the English tables read as metric values: what
    #    matters for parity is that every conversion site is exercised (energy-balance snow with ice, monthly CEPSC / UZSN /
    #    RETSC, frozen-ground factor, PWATER without SNOW) ------------------------------------------------------------
    t = test10.perlnd()
    t["ACTIVITY"].update({"PSTEMP": 0, "PWTGAS": 0})
    add("p001_metric", "PERLND", t, "perlnd", drop=("CLOUD",), units=2)
    t = test10.perlnd()
    t["ACTIVITY"].update({"PSTEMP": 0, "PWTGAS": 0})
    t["SNOW"]["STATES"].update({"PACKF": 40.0, "PACKI": 6.0, "PACKW": 2.0, "PAKTMP": -3.0, "COVINX": 20.0})
    t["SNOW"]["PARAMETERS"].update({"TSNOW": 1.0, "TBASE": 0.0, "COVIND": 30.0, "MGMELT": 0.5, "MELEV": 300.0})
    t["PWATER"]["PARAMETERS"].update({"PETMAX": 4.0, "PETMIN": 1.5, "KVARY": 0.02, "FZG": 0.04})
    add("p_metric_pack", "PERLND", t, "perlnd", units=2, rain_mult=2.0)
    ti = test10.implnd()
    ti["ACTIVITY"].update({"SOLIDS": 0, "IWTGAS": 0})
    ti["IWATER"]["PARAMETERS"].update({"VRSFG": 1})
    ti["IWATER"]["MONTHLY_RETSC"] = monthly([.05, .05, .06, .07, .08, .1, .1, .1, .08, .07, .06, .05])
    add("i001_metric", "IMPLND", ti, "implnd", units=2)
    fp = {"ACTIVITY": {"PWATER": 1}, "PWATER": copy.deepcopy(FLOWTEST_PERLND)}
    add("flowtest_pwater_metric", "PERLND", fp, "flowtest", steps=8760, year=2018, units=2)
    # sub-hourly metric run: the time conversions (MGMELT, KMELT, INFILT * delt60) come before the unit conversions
    t = test10.perlnd()
    t["ACTIVITY"].update({"PSTEMP": 0, "PWTGAS": 0})
    add("p001_metric_delt30", "PERLND", t, "perlnd", steps=2 * 24 * 100, delt=30, units=2)

    # -- metric units in the other sections (SEDMNT.py:113-117,258-262; PSTEMP.py:80-126,136-149,205-215; PWTGAS.py:90-94,167-174,
    #    310-350; SOLIDS.py:84-88,143-145; IWTGAS.py:80-82,111-114,160-163,179-182): again the English tables read as metric
    #    values.  The full CBP chain (TSOPFG=1 tables, method-1 washoff, lateral inflows with concentrations); the TSOPFG=0 /
    #    VSIVFG=2 / method-2 variant WITHOUT a PSTEMP initial-state table (the 16 deg C defaults of PSTEMP.py:80-88); the IMPLND
    #    chain with monthly accumulation and regression tables; SOLIDS / IWTGAS on external runoff series ----------------
    t = _cbp_tables()
    t["PWTGAS"]["PARAMETERS"].update({"SLIFAC": 0.5, "ILIFAC": 0.3, "ALIFAC": 0.2})
    t["PSTEMP"]["STATES"] = {"AIRTC": 5.0, "SLTMP": 4.0, "ULTMP": 6.0, "LGTMP": 8.0}
    add("cbp_fullchain_metric", "PERLND", t, "perlnd", gatmp=2.0, units=2,
        lateral=("SURLI", "UZLI", "IFWLI", "LZLI", "AGWLI", "SLSED"),
        lateral_conc={"SLITMP": (2., 20.), "SLIDOX": (4., 12.), "SLICO2": (0., .2), "ILITMP": (2., 20.),
                      "ILIDOX": (4., 12.), "ILICO2": (0., .2), "ALITMP": (4., 15.), "ALIDOX": (3., 9.),
                      "ALICO2": (0., .2)})
    t = _cbp_tables()
    t["PWATER"]["PARAMETERS"].update({"RTOPFG": 0, "UZFG": 1, "VLEFG": 2, "KVARY": 0.0, "BASETP": 0.15})
    t["SEDMNT"]["PARAMETERS"].update({"SDOPFG": 0, "VSIVFG": 2, "KGER": 0.3, "JGER": 1.5})
    t["PSTEMP"]["PARAMETERS"].update({"TSOPFG": 0, "SLTVFG": 0, "ULTVFG": 0, "LGTVFG": 0, "ASLT": 14.5,
                                      "BSLT": 0.365, "ULTP1": 0.1, "ULTP2": 4.0, "LGTP1": 0.05, "LGTP2": 6.0})
    for grp in ("PARAMETERS", "STATES"):  # no initial soil temperatures: 16 deg C (PSTEMP.py:80-88)
        for k in ("AIRTC", "SLTMP", "ULTMP", "LGTMP"):
            t["PSTEMP"].get(grp, {}).pop(k, None)
    t["PWTGAS"]["PARAMETERS"].update({"IDVFG": 0, "GDVFG": 0, "ICVFG": 1, "GCVFG": 1, "IDOXP": 6.0, "ADOXP": 5.0})
    t["PWTGAS"]["MONTHLY_ICO2P"] = monthly([.05, .05, .06, .07, .08, .1, .12, .12, .1, .08, .06, .05])
    t["PWTGAS"]["MONTHLY_ACO2P"] = monthly([.04, .04, .05, .05, .06, .07, .08, .08, .07, .06, .05, .04])
    add("chain_uzfg_newton_metric", "PERLND", t, "perlnd", drop=("CLOUD",), gatmp=-1.5, rain_mult=3.0, units=2)
    t = test10.implnd()
    t["ACTIVITY"].update({"ATEMP": 1, "IWTGAS": 1})
    t["ATEMP"] = {"PARAMETERS": {"ELDAT": 120.0, "AIRTMP": 30.0}}
    t["SOLIDS"]["PARAMETERS"].update({"SDOPFG": 1, "VASDFG": 1, "VRSDFG": 1, "KEIM": 0.3, "JEIM": 1.4})
    t["SOLIDS"]["MONTHLY_ACCSDP"] = monthly([.004, .004, .006, .01, .012, .014, .014, .012, .01, .008, .006, .004])
    t["SOLIDS"]["MONTHLY_REMSDP"] = monthly([.03, .03, .04, .05, .06, .07, .07, .06, .05, .04, .03, .03])
    t["IWTGAS"]["PARAMETERS"].update({"WTFVFG": 1, "SLIFAC": 0.4, "ELEV": 380.0})
    t["IWTGAS"]["STATES"] = {"SOTMP": 15.0, "SODOX": 0.0, "SOCO2": 0.0}
    t["IWTGAS"]["MONTHLY_AWTF"] = monthly([1., 2., 3., 7., 11., 16., 19., 18., 14., 9., 4., 2.])
    t["IWTGAS"]["MONTHLY_BWTF"] = monthly([.3, .3, .35, .4, .45, .5, .5, .5, .45, .4, .35, .3])
    add("i_chain_metric", "IMPLND", t, "implnd", gatmp=1.0, drop=("CLOUD",), rain_mult=2.5, lateral=("SURLI",),
        lateral_conc={"SLITMP": (2., 20.), "SLIDOX": (4., 12.)}, units=2)
    t = test10.implnd()
    t["ACTIVITY"].update({"SNOW": 0, "IWATER": 0, "IWTGAS": 1})
    t["IWTGAS"]["STATES"] = {"SOTMP": 15.0, "SODOX": 0.0, "SOCO2": 0.0}
    add("i_solids_iwtgas_alone_metric", "IMPLND", t, "implnd", steps=24 * 120, runoff_inputs=True, units=2)
    return C


# tests/ipwater/data/uci_data.json (numeric entries the kernels read; the rest of that file is unused by them)
FLOWTEST_IMPLND = {
    "PARAMETERS": {"CSNOFG": 0, "RTOPFG": 0, "VRSFG": 0, "VNNFG": 0, "RTLIFG": 0, "LSUR": 100.0, "SLSUR": 0.02,
                   "NSUR": 0.05, "RETSC": 0.15, "PETMAX": 40.0, "PETMIN": 35.0},
    "STATES": {"RETS": 0.01, "SURS": 0.01},
}
FLOWTEST_PERLND = {
    "PARAMETERS": {"CSNOFG": 0, "RTOPFG": 1, "UZFG": 1, "VCSFG": 1, "VUZFG": 0, "VNNFG": 0, "VIFWFG": 0, "VIRCFG": 0,
                   "VLEFG": 0, "IFFCFG": 1, "HWTFG": 0, "IRRGFG": 0, "IFRDFG": 0, "FOREST": 0.0, "LZSN": 7.35,
                   "INFILT": 0.079, "LSUR": 50.0, "SLSUR": 0.1548, "KVARY": 1.0, "AGWRC": 0.993, "PETMAX": 40.0,
                   "PETMIN": 35.0, "INFEXP": 2.0, "INFILD": 2.0, "DEEPFR": 0.02, "BASETP": 0.22, "AGWETP": 0.0,
                   "CEPSC": 0.01, "UZSN": 0.85, "NSUR": 0.4, "INTFW": 3.0, "IRC": 0.7, "LZETP": 0.73, "FZG": 1.0,
                   "FZGL": 0.1},
    "STATES": {"CEPS": 0.0462, "SURS": 0.0, "UZS": 1.6075, "IFWS": 0.0996, "LZS": 8.9629, "AGWS": 2.8649,
               "GWVS": 1.9555},
    "MONTHLY_CEPSC": monthly([0.09, 0.1, 0.1, 0.14, 0.17, 0.18, 0.18, 0.18, 0.17, 0.15, 0.12, 0.1]),
}
