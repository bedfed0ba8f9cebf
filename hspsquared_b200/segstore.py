"""Batched segment store: packs land segments into the structure-of-arrays buffers the kernels read.

Host-side restatement of what the reference's activity wrappers do per segment before they enter their @njit core
(SNOW.py:35-79, PWATER.py:35-93, IWATER.py:31-66, SEDMNT.py:24-43, PSTEMP.py:25-53, PWTGAS.py:32-56, ATEMP.py:19-24,
utilities.py:58-79 make_numba_dict) and of the cores' pre-loop initialisation (SNOW.py:102-232, PWATER.py:159-260,
IWATER.py:128-143, SEDMNT.py:63-145, PSTEMP.py:76-128, PWTGAS.py:73-95):

  effective(operation, tables)   one segment's UCI tables  -> flat effective values (defaults, flag rules, which
                                 parameters follow a monthly table);
  SegmentStore.build(...)        arrays of effective values -> par[NPAR][n], 64-bit flag word per segment,
                                 monthly tables [12][n], full initial state[NSTATE][n]  (vectorised numpy: every
                                 expression is the reference's, elementwise IEEE double, hence bit-identical);
  forcing_fill(...)              replays the wrappers' NaN / zero / -1e30 fills for absent input series.

Unsupported (refused loudly, like the reference or because the reference itself cannot run them):
PWATER HWTFG=1 (PWATER.py:107-110), IFRDFG outside {0,1}, a broadcast parameter supplied as an
external time series (SNOW.py:49-51, PWATER.py:51-53).
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib

NAN = float("nan")

PERLND_ORDER = ("ATEMP", "SNOW", "PWATER", "SEDMNT", "PSTEMP", "PWTGAS")
IMPLND_ORDER = ("ATEMP", "SNOW", "IWATER", "SOLIDS", "IWTGAS")

# raw (pre-derivation) per-segment values collected by effective(); derived columns are made in build()
RAW_DEFAULT = {
    "ELDAT": 0.0,
    "MELEV": 0.0, "SHADE": 0.0, "SNOWCF": 1.0, "COVIND": 1.0, "KMELT": 0.0, "TBASE": 32.0, "RDCSN": 0.15,
    "TSNOW": 32.0, "SNOEVP": 0.1, "CCFACT": 1.0, "MWATER": 0.03, "MGMELT": 0.01,
    "FOREST": 0.0, "LZSN": 1.0, "INFILT": 0.0, "LSUR": 1.0, "SLSUR": 1.0, "KVARY": 0.0, "AGWRC": 0.0, "PETMAX": 40.0,
    "PETMIN": 35.0, "INFEXP": 2.0, "INFILD": 2.0, "DEEPFR": 0.0, "BASETP": 0.0, "AGWETP": 0.0, "CEPSC": 0.0,
    "UZSN": 1.0, "NSUR": 0.1, "INTFW": 0.0, "IRC": 0.5, "LZETP": 0.0, "FZG": 1.0, "FZGL": 0.1,
    "SMPF": 1.0, "KRER": 0.0, "JRER": 1.0, "AFFIX": 0.0, "COVER": 0.0, "NVSI": 0.0, "KSER": 0.0, "JSER": 1.0,
    "KGER": 0.0, "JGER": 1.0,
    "ASLT": 32.0, "BSLT": 1.0, "ULTP1": 0.0, "ULTP2": 0.0, "LGTP1": 0.0, "LGTP2": 0.0,
    "ELEV": 0.0, "IDOXP": 0.0, "ICO2P": 0.0, "ADOXP": 0.0, "ACO2P": 0.0, "SLIFAC": 0.0, "ILIFAC": 0.0, "ALIFAC": 0.0,
    "RETSC": 0.0,
    "KEIM": 0.0, "JEIM": -999.0, "ACCSDP": 0.0, "REMSDP": 0.0, "AWTF": 32.0, "BWTF": 1.0,  # ParseTable.csv rows 1291-1329
    # columns of PQUAL / IQUAL rows (qualstore.py); unused by land-segment batches
    "Q_WSFAC": 0.0, "Q_POTFW": 0.0, "Q_POTFS": 0.0, "Q_ACQOP": 0.0, "Q_REMQOP": 0.0, "Q_IOQC": 0.0, "Q_AOQC": 0.0,
    "Q_ADFX": 0.0, "Q_ADCN": 0.0,
}
INIT_DEFAULT = {
    "PACKF": 0.0, "PACKI": 0.0, "PACKW": 0.0, "RDENPF": 0.2, "DULL": 400.0, "PAKTMP": 32.0, "COVINX": 0.01,
    "XLNMLT": 0.0, "SKYCLR": 1.0,
    "CEPS": 0.0, "SURS": 0.0, "UZS": 0.001, "IFWS": 0.0, "LZS": 0.001, "AGWS": 0.0, "GWVS": 0.0,
    "RETS": 0.001,
    "SLDS": 0.0,
    "DETS": 0.0, "PST_SLTMP": NAN, "PST_ULTMP": NAN, "PST_LGTMP": NAN,  # NaN = absent: 60 deg F / 16 deg C (PSTEMP.py:80-88)
}


def make_ui(uci):
    """utilities.py:58-79: merge FLAGS/PARAMETERS/STATES keeping exact Python int/float values only."""
    ui = {}
    for name in ("FLAGS", "PARAMETERS", "STATES"):
        for key, value in (uci.get(name) or {}).items():
            if type(value) in (int, float):
                ui[key] = float(value)
    return ui


def _need(ui, key, where):
    if key not in ui:
        raise KeyError("%s: required value %r is missing (the reference raises KeyError here too)" % (where, key))
    return ui[key]


def _table(uci, key):
    t = uci[key]
    vals = list(t.values()) if isinstance(t, dict) else list(t)
    if len(vals) != 12:
        raise ValueError("%s must hold 12 monthly values" % key)
    return [float(v) for v in vals]


def _initm(uci, flag, key):
    """utilities.py:205-211: the table is used only when the flag is truthy AND the table exists."""
    return bool(flag) and key in uci


def effective(operation, tables, pw_icefg=None, airtfg=None, ext_names=()):
    """One segment: {'act': mask, 'raw': {...}, 'flag': {...}, 'mon': {table: [12]}, 'init': {...}}.

    `tables` = {'ACTIVITY': {name: 0/1}, '<ACTIVITY NAME>': uci[(operation, activity, segment)] ...}.
    pw_icefg: ICEFG as PWATER sees it via siminfo (SNOW.py:60-63, PWATER.py:87); default = the segment's own SNOW
    ICEFG when SNOW is active, else 0.  airtfg: main.py:141-143 injection; default = ACTIVITY['ATEMP'].
    """
    order = PERLND_ORDER if operation == "PERLND" else IMPLND_ORDER
    flags_act = tables.get("ACTIVITY", {})
    act = 0
    for a in order:
        if flags_act.get(a, 0) and a in tables:
            act |= _lib.ACT_BITS[a]
    raw, flag, mon, init = {}, {}, {}, {}
    ext = set(ext_names)

    def injected_csnofg(ui, where):
        """main.py:138-146 copies PWATER's CSNOFG (as rewritten by PWATER.py:89-93) into the SEDMNT / PWTGAS tables."""
        if "CSNOFG" in ui:
            return int(ui["CSNOFG"])
        if "PWATER" in tables:
            pw = make_ui(tables["PWATER"])
            return int(pw["CSNOFG"]) if "CSNOFG" in pw else 0
        raise KeyError("%s: CSNOFG is missing (main.py injects it from the PWATER table)" % where)

    def no_ext(names, where):
        bad = [n for n in names if n in ext]
        if bad:
            raise NotImplementedError("%s: parameter(s) %s supplied as external time series are not supported" % (where, bad))

    if act & _lib.ACT_BITS["ATEMP"]:
        ui = make_ui(tables["ATEMP"])
        raw["ELDAT"] = _need(ui, "ELDAT", "ATEMP")

    if act & _lib.ACT_BITS["SNOW"]:
        uci = tables["SNOW"]
        ui = make_ui(uci)
        u = uci["PARAMETERS"]
        no_ext(("CCFACT", "COVIND", "MGMELT", "MWATER", "SHADE", "SNOEVP", "SNOWCF", "KMELT"), "SNOW")
        for n in ("CCFACT", "COVIND", "MGMELT", "MWATER", "SHADE", "SNOEVP", "SNOWCF"):
            raw[n] = float(_need(u, n, "SNOW"))
        for n in ("MELEV", "RDCSN", "TBASE", "TSNOW"):
            raw[n] = _need(ui, n, "SNOW")
        raw["KMELT"] = float(_need(u, "KMELT", "SNOW"))
        flag["ICEFG"] = int(ui.get("ICEFG", 0))
        flag["SNOPFG"] = int(ui.get("SNOPFG", 0))
        vkmfg = (uci.get("FLAGS") or {}).get("VKMFG", 0)
        if _initm(uci, vkmfg, "MONTHLY_KMELT"):
            flag["M_KMELT"] = 1
            mon["KMELT"] = _table(uci, "MONTHLY_KMELT")
        for n in ("PACKF", "PACKI", "PACKW", "PAKTMP", "RDENPF", "SKYCLR", "COVINX", "DULL", "XLNMLT"):
            init[n] = _need(ui, n, "SNOW")

    if act & _lib.ACT_BITS["PWATER"]:
        uci = tables["PWATER"]
        ui = make_ui(uci)
        u = uci["PARAMETERS"]
        if int(ui.get("HWTFG", 0)):
            raise NotImplementedError("PWATER HWTFG=1: not implemented in the reference either (PWATER.py:107-110)")
        no_ext(("AGWRC", "DEEPFR", "INFILT", "KVARY", "LZSN", "PETMIN", "PETMAX"), "PWATER")
        for n in ("AGWRC", "DEEPFR", "INFILT", "KVARY", "LZSN", "PETMIN", "PETMAX"):
            raw[n] = float(_need(u, n, "PWATER"))
        for n in ("AGWETP", "BASETP", "INFEXP", "INFILD", "LSUR", "SLSUR", "FOREST"):
            raw[n] = _need(ui, n, "PWATER")
        if "CSNOFG" in ui:  # PWATER.py:170-176
            csnofg, iffcfg, ifrdfg = int(ui["CSNOFG"]), int(ui["IFFCFG"]), int(ui["IFRDFG"])
            rtopfg, uzfg, vlefg = int(ui["RTOPFG"]), int(ui["UZFG"]), int(ui["VLEFG"])
        else:
            csnofg, iffcfg, ifrdfg, rtopfg, uzfg, vlefg = 0, 1, 0, 0, 0, 0
        if ifrdfg not in (0, 1):
            raise NotImplementedError("IFRDFG must be 0 or 1")
        flag.update(CSNOFG=csnofg, IFFCFG2=int(iffcfg == 2), IFRDFG=ifrdfg, RTOP1=int(rtopfg == 1), UZFG=int(bool(uzfg)),
                    VLE_GE2=int(vlefg >= 2))
        if pw_icefg is None:
            pw_icefg = flag.get("ICEFG", 0) if act & _lib.ACT_BITS["SNOW"] else 0
        flag["PW_ICEFG"] = int(bool(int(pw_icefg)))
        raw["FZG"] = ui.get("FZG", 1.0) if flag["PW_ICEFG"] else 0.0
        raw["FZGL"] = ui.get("FZGL", 0.1) if flag["PW_ICEFG"] else 0.0
        if flag["PW_ICEFG"]:
            _need(ui, "FZG", "PWATER")
        for n in ("LZETP", "CEPSC", "INTFW", "IRC", "NSUR", "UZSN"):
            raw[n] = float(_need(u, n, "PWATER"))
        if "VLEFG" in u:  # PWATER.py:57-65
            rules = {"LZETP": (u["VLEFG"] == 1) or (u["VLEFG"] == 3), "CEPSC": u["VCSFG"], "INTFW": u["VIFWFG"],
                     "IRC": u["VIRCFG"], "NSUR": u["VNNFG"], "UZSN": u["VUZFG"]}
            for n, f in rules.items():
                if _initm(uci, f, "MONTHLY_" + n):
                    flag["M_" + n] = 1
                    mon[n] = _table(uci, "MONTHLY_" + n)
        for n in ("CEPS", "SURS", "UZS", "IFWS", "LZS", "AGWS", "GWVS"):
            init[n] = _need(ui, n, "PWATER")

    if act & _lib.ACT_BITS["IWATER"]:
        uci = tables["IWATER"]
        ui = make_ui(uci)
        u = uci["PARAMETERS"]
        no_ext(("PETMAX", "PETMIN"), "IWATER")
        raw["PETMAX"] = float(_need(u, "PETMAX", "IWATER"))
        raw["PETMIN"] = float(_need(u, "PETMIN", "IWATER"))
        raw["LSUR"] = _need(ui, "LSUR", "IWATER")
        raw["SLSUR"] = _need(ui, "SLSUR", "IWATER")
        raw["RETSC"] = float(_need(u, "RETSC", "IWATER"))
        raw["NSUR"] = float(_need(u, "NSUR", "IWATER"))
        flag.update(CSNOFG=int(bool(int(ui.get("CSNOFG", 0)))), RTOP1=int(bool(int(ui.get("RTOPFG", 0)))),
                    RTLIFG=int(bool(int(ui.get("RTLIFG", 0)))))
        if "VRSFG" in u:  # IWATER.py:48-54
            if _initm(uci, u["VRSFG"], "MONTHLY_RETSC"):
                flag["M_RETSC"] = 1
                mon["RETSC"] = _table(uci, "MONTHLY_RETSC")
            if _initm(uci, u["VNNFG"], "MONTHLY_NSUR"):
                flag["M_NSUR"] = 1
                mon["NSUR"] = _table(uci, "MONTHLY_NSUR")
        init["RETS"] = _need(ui, "RETS", "IWATER")
        init["SURS"] = _need(ui, "SURS", "IWATER")

    if act & _lib.ACT_BITS["SOLIDS"]:  # SOLIDS.py:19-49
        uci = tables["SOLIDS"]
        ui = make_ui(uci)
        u = uci["PARAMETERS"]
        for n in ("KEIM", "JEIM"):
            raw[n] = _need(ui, n, "SOLIDS")
        for n in ("ACCSDP", "REMSDP"):
            raw[n] = float(_need(u, n, "SOLIDS"))
        init["SLDS"] = _need(ui, "SLDS", "SOLIDS")
        flag["SLD_SDOP1"] = int(ui.get("SDOPFG", 0.0) == 1)
        for fl, n in (("VASDFG", "ACCSDP"), ("VRSDFG", "REMSDP")):
            if fl in u and _initm(uci, u[fl], "MONTHLY_" + n):
                flag["M_" + n] = 1
                mon[n] = _table(uci, "MONTHLY_" + n)

    if act & _lib.ACT_BITS["IWTGAS"]:  # IWTGAS.py:33-63, 76-86
        uci = tables["IWTGAS"]
        ui = make_ui(uci)
        u = uci["PARAMETERS"]
        for n in ("SOTMP", "SODOX", "SOCO2"):
            _need(ui, n, "IWTGAS")  # initial values the loop overwrites, but their absence is a KeyError in the reference
        for n in ("SLIFAC", "ELEV"):
            raw[n] = _need(ui, n, "IWTGAS")
        for n in ("AWTF", "BWTF"):
            raw[n] = float(_need(u, n, "IWTGAS"))
            if "WTFVFG" in u and _initm(uci, u["WTFVFG"], "MONTHLY_" + n):
                flag["M_" + n] = 1
                mon[n] = _table(uci, "MONTHLY_" + n)

    if act & _lib.ACT_BITS["SEDMNT"]:
        uci = tables["SEDMNT"]
        ui = make_ui(uci)
        u = uci["PARAMETERS"]
        for n in ("SMPF", "KRER", "JRER", "AFFIX", "NVSI", "KSER", "JSER", "KGER", "JGER"):
            raw[n] = _need(ui, n, "SEDMNT")
        raw["COVER"] = float(_need(u, "COVER", "SEDMNT"))
        vsivfg = ui.get("VSIVFG", 0.0)
        flag.update(SED_CSNO1=int(injected_csnofg(ui, "SEDMNT") == 1), SDOPFG=int(ui.get("SDOPFG", 0.0) != 0.0),
                    VSIV_EQ2=int(vsivfg == 2), VSIV_LE2=int(vsivfg <= 2))
        if "CRVFG" in u and _initm(uci, u["CRVFG"], "MONTHLY_COVER"):
            flag["M_COVER"] = 1
            mon["COVER"] = _table(uci, "MONTHLY_COVER")
        if "VSIVFG" in u and _initm(uci, u["VSIVFG"], "MONTHLY_NVSI"):
            flag["M_NVSI"] = 1
            mon["NVSI"] = _table(uci, "MONTHLY_NVSI")
        init["DETS"] = _need(ui, "DETS", "SEDMNT")

    if act & _lib.ACT_BITS["PSTEMP"]:
        uci = tables["PSTEMP"]
        ui = make_ui(uci)
        u = uci["PARAMETERS"]
        for n in ("ASLT", "BSLT", "ULTP1", "ULTP2", "LGTP1", "LGTP2"):
            raw[n] = float(_need(u, n, "PSTEMP"))
        tsopfg = ui.get("TSOPFG", 0.0)
        if airtfg is None:
            airtfg = ui["AIRTFG"] if "AIRTFG" in ui else flags_act.get("ATEMP", 0)
        flag.update(TSOP1=int(tsopfg == 1), TSOP0=int(tsopfg == 0), AIRTFG=int(int(airtfg) != 0))
        for fl, names in (("SLTVFG", ("ASLT", "BSLT")), ("ULTVFG", ("ULTP1", "ULTP2")), ("LGTVFG", ("LGTP1", "LGTP2"))):
            for n in names:
                if fl in u and _initm(uci, u[fl], "MONTHLY_" + n):
                    flag["M_" + n] = 1
                    mon[n] = _table(uci, "MONTHLY_" + n)
        init["PST_SLTMP"] = ui.get("SLTMP", NAN)  # absent: the unit system's default, SegmentStore.build
        init["PST_ULTMP"] = ui.get("ULTMP", NAN)
        init["PST_LGTMP"] = ui.get("LGTMP", NAN)

    if act & _lib.ACT_BITS["PWTGAS"]:
        uci = tables["PWTGAS"]
        ui = make_ui(uci)
        u = uci["PARAMETERS"]
        for n in ("SOTMP", "IOTMP", "AOTMP", "SODOX", "SOCO2", "IODOX", "IOCO2", "AODOX", "AOCO2", "SDLFAC"):
            _need(ui, n, "PWTGAS")  # read (and unused) by the reference, but their absence is a KeyError there
        for n in ("SLIFAC", "ILIFAC", "ALIFAC", "ELEV"):
            raw[n] = _need(ui, n, "PWTGAS")
        for n in ("IDOXP", "ICO2P", "ADOXP", "ACO2P"):
            raw[n] = float(_need(u, n, "PWTGAS"))
        flag.update(PWG_CSNO=int(bool(injected_csnofg(ui, "PWTGAS"))), GDVFG=int(ui.get("GDVFG", 0.0) != 0.0))
        for fl, n in (("IDVFG", "IDOXP"), ("ICVFG", "ICO2P"), ("GDVFG", "ADOXP"), ("GCVFG", "ACO2P")):
            if fl in u and _initm(uci, u[fl], "MONTHLY_" + n):
                flag["M_" + n] = 1
                mon[n] = _table(uci, "MONTHLY_" + n)

    return {"act": act, "raw": raw, "flag": flag, "mon": mon, "init": init}


def stack(effs):
    """List of effective() records (same activity mask) -> arrays keyed like one record."""
    n = len(effs)
    act = effs[0]["act"]
    if any(e["act"] != act for e in effs):
        raise ValueError("segments of one batch must share the activity mask")
    raw = {k: np.array([e["raw"].get(k, RAW_DEFAULT[k]) for e in effs], dtype=np.float64)
           for k in set().union(*[e["raw"].keys() for e in effs])}
    flag = {k: np.array([e["flag"].get(k, 0) for e in effs], dtype=np.uint64)
            for k in set().union(*[e["flag"].keys() for e in effs])}
    init = {k: np.array([e["init"].get(k, INIT_DEFAULT[k]) for e in effs], dtype=np.float64)
            for k in set().union(*[e["init"].keys() for e in effs])}
    mon = {}
    for k in set().union(*[e["mon"].keys() for e in effs]):
        t = np.zeros((12, n))
        for i, e in enumerate(effs):
            if k in e["mon"]:
                t[:, i] = e["mon"][k]
        mon[k] = t
    return {"act": act, "raw": raw, "flag": flag, "mon": mon, "init": init, "n": n}


def broadcast(eff, n):
    """One effective() record replicated to n segments (base of the synthetic ensembles)."""
    return {
        "act": eff["act"], "n": n,
        "raw": {k: np.full(n, float(v)) for k, v in eff["raw"].items()},
        "flag": {k: np.full(n, int(v), dtype=np.uint64) for k, v in eff["flag"].items()},
        "init": {k: np.full(n, float(v)) for k, v in eff["init"].items()},
        "mon": {k: np.repeat(np.asarray(v, dtype=np.float64)[:, None], n, axis=1) for k, v in eff["mon"].items()},
    }


def _vpow(base, expo):
    """Elementwise glibc pow (what Numba's array ** scalar reaches); numpy's SIMD pow is not bit-identical."""
    expo = float(expo)
    return np.fromiter((math.pow(float(b), expo) for b in base), dtype=np.float64, count=len(base))


class SegmentStore:
    """SoA host image of one batch: par [NPAR][n], flags [n], monthly {id: [12][n]}, state [NSTATE][n]."""

    units = 1  # 1 English, 2 metric (siminfo['units'])

    def __init__(self, operation, act, n, delt):
        self.operation = operation
        self.act = int(act)
        self.n = int(n)
        self.delt = float(delt)
        self.P, self.S = _lib.index("params"), _lib.index("states")
        self.FB, self.M = _lib.index("flags"), _lib.index("monthly")
        self.par = np.zeros((len(self.P), n), dtype=np.float64)
        self.flags = np.zeros(n, dtype=np.uint64)
        self.state = np.zeros((len(self.S), n), dtype=np.float64)
        self.monthly = {}

    @classmethod
    def from_tables(cls, operation, tables_list, calendar, pw_icefg=None, airtfg=None, ext_names=(), units=1):
        effs = [effective(operation, t, pw_icefg, airtfg, ext_names) for t in tables_list]
        return cls.build(operation, stack(effs), calendar, units=units)

    # metric units (siminfo['units'] == 2): what the reference converts on entry of each core, in its expression order
    # (`x * 9./5. + 32.` is ((x * 9.) / 5.) + 32.).  The kernels work in English units; the per-interval conversions
    # (parameter series after monthly interpolation, the SNOW -> water link, the stored series) are theirs (HSP2G_OPT_METRIC).
    METRIC_RAW = {
        "SNOW": {"MELEV": lambda x: x * 3.28084, "TBASE": lambda x: (x * 9. / 5.) + 32., "TSNOW": lambda x: (x * 9. / 5.) + 32.,
                 "COVIND": lambda x: x / 25.4},  # SNOW.py:120,131-133,157-158; MGMELT below (after its time conversion)
        "PWATER": {"LSUR": lambda x: x * 3.28, "KVARY": lambda x: x / 0.0394, "LZSN": lambda x: x * 0.0394,
                   "PETMAX": lambda x: (x * 9. / 5.) + 32., "PETMIN": lambda x: (x * 9. / 5.) + 32.,
                   "FZG": lambda x: x * 0.0394},  # PWATER.py:194,206-207,240-243; INFILT below; CEPSC / UZSN on the device
        "IWATER": {"LSUR": lambda x: x * 3.28, "PETMAX": lambda x: (x * 9. / 5.) + 32., "PETMIN": lambda x: (x * 9. / 5.) + 32.},
        "PWTGAS": {"ELEV": lambda x: x * 3.281},  # PWTGAS.py:94 (before ELEVGC); SOTMP / IOTMP / AOTMP are overwritten before use
        "IWTGAS": {"ELEV": lambda x: x * 3.281},  # IWTGAS.py:82
    }
    METRIC_INIT = {
        "SNOW": {"PACKI": lambda x: x / 25.4, "PACKW": lambda x: x / 25.4, "COVINX": lambda x: x / 25.4,
                 "PAKTMP": lambda x: (x * 9. / 5.) + 32.},  # SNOW.py:114-119; PACKF below (the ice is added first)
        "PWATER": {k: (lambda x: x * 0.0394) for k in ("CEPS", "SURS", "UZS", "IFWS", "LZS", "AGWS", "GWVS")},  # PWATER.py:195-201
        "IWATER": {"RETS": lambda x: x * 0.0394, "SURS": lambda x: x * 0.0394},  # IWATER.py:131-133
        "SEDMNT": {"DETS": lambda x: x * 1.10231 / 2.471},  # SEDMNT.py:117; the NVSI series on the device
        "SOLIDS": {"SLDS": lambda x: x * 1.10231 / 2.471},  # SOLIDS.py:88; the ACCSDP series on the device
    }

    @classmethod
    def build(cls, operation, arrs, calendar, units=1):
        n, act = arrs["n"], arrs["act"]
        delt = float(calendar.delt)
        delt60 = delt / 60.0
        st = cls(operation, act, n, delt)
        raw, flag, init, mon = arrs["raw"], arrs["flag"], arrs["init"], arrs["mon"]
        B = _lib.ACT_BITS
        if units not in (1, 2):
            raise ValueError("units must be 1 (English) or 2 (metric)")
        st.units = int(units)
        metric = units == 2
        conv_raw, conv_init = {}, {}
        if metric:
            # every section of the path (ATEMP and the quality constituents have no unit handling in the reference); PSTEMP's
            # tables and states are deg C as they are, below
            for sec in ("SNOW", "PWATER", "IWATER", "SEDMNT", "PWTGAS", "SOLIDS", "IWTGAS"):
                if act & B[sec]:
                    conv_raw.update(cls.METRIC_RAW.get(sec, {}))
                    conv_init.update(cls.METRIC_INIT.get(sec, {}))

        def R(name):
            v = raw.get(name, np.full(n, RAW_DEFAULT[name]))
            return conv_raw[name](v) if name in conv_raw else v

        def I(name):
            v = init.get(name, np.full(n, INIT_DEFAULT[name]))
            return conv_init[name](v) if name in conv_init else v

        # ---- plain parameter columns
        derived = {"MGMELT_IVL", "INFILT_IVL", "KGW", "ELEVGC", "MELEV_FAC", "SNOEVP_K", "CCFACT_K"}
        for name, row in st.P.items():
            if name not in derived:
                st.par[row] = R(name)
        st.par[st.P["MGMELT_IVL"]] = R("MGMELT") * delt / 1440.0  # SNOW.py:146
        st.par[st.P["INFILT_IVL"]] = R("INFILT") * delt60  # PWATER.py:215
        # per-segment constant heads of SNOW expressions (the first operations of the reference's left-to-right evaluation)
        st.par[st.P["MELEV_FAC"]] = 1.0 - 0.3 * R("MELEV") / 10000.0  # SNOW.py:360  reltmp * (1.0 - 0.3 * melev / 10000.0) * factr
        st.par[st.P["SNOEVP_K"]] = R("SNOEVP") * 0.0002  # SNOW.py:336  snoevp * 0.0002 * winmov * ...
        st.par[st.P["CCFACT_K"]] = R("CCFACT") * 0.00026  # SNOW.py:354  ccfact * 0.00026 * winmov
        if metric:
            st.par[st.P["MGMELT_IVL"]] = st.par[st.P["MGMELT_IVL"]] / 25.4  # SNOW.py:159
            st.par[st.P["INFILT_IVL"]] = st.par[st.P["INFILT_IVL"]] * 0.0394  # PWATER.py:239
        if act & B["PWATER"]:
            st.par[st.P["KGW"]] = 1.0 - _vpow(R("AGWRC"), delt60 / 24.0)  # PWATER.py:247
        if act & (B["PWTGAS"] | B["IWTGAS"]):
            st.par[st.P["ELEVGC"]] = _vpow((288.0 - 0.00198 * R("ELEV")) / 288.0, 5.256)  # PWTGAS.py:95, IWTGAS.py:84

        # ---- flag word and monthly tables
        for name, v in flag.items():
            st.flags |= (v.astype(np.uint64) & np.uint64(1)) << np.uint64(st.FB[name])
        for name, tab in mon.items():
            st.monthly[st.M[name]] = np.ascontiguousarray(tab, dtype=np.float64)

        def daily0(name):
            """value at interval 0 of a parameter that may follow a monthly table (initm + dayval at step 0)."""
            v = R(name).copy()
            if name in mon:
                on = ((st.flags >> np.uint64(st.FB["M_" + name])) & np.uint64(1)).astype(bool)
                m0, m1, xm, dxm, _ = calendar.day_interp()
                y0, y1 = mon[name][m0[0]], mon[name][m1[0]]
                with np.errstate(invalid="ignore", divide="ignore"):
                    t = ((y1 - y0) / dxm[0]) * xm[0] + y0
                t = y0 if xm[0] == 0.0 else t
                v[on] = t[on]
            return v

        S = st.S
        s = st.state
        if act & B["SNOW"]:  # SNOW.py:102-232
            covinx, dull = I("COVINX").copy(), I("DULL").copy()
            packf = init.get("PACKF", np.full(n, INIT_DEFAULT["PACKF"])) + init.get("PACKI", np.full(n, INIT_DEFAULT["PACKI"]))
            if metric:
                packf = packf / 25.4  # SNOW.py:110,115: the ice is added in the user's units, then the sum is converted
            packi, packw, paktmp = I("PACKI").copy(), I("PACKW").copy(), I("PAKTMP").copy()
            rdenpf = I("RDENPF").copy()
            covind0 = R("COVIND")
            nopack = (packf + packw) <= 1.0e-5
            with np.errstate(invalid="ignore", divide="ignore"):
                covinx = np.where(nopack | (covinx < 1.0e-5), 0.1 * covind0, covinx)
                pdepth = np.where(nopack, 0.0, packf / rdenpf)
                snocov = np.where(nopack, 0.0, np.where(packf < covinx, packf / covinx, 1.0))
                neghts = np.where(nopack, 0.0, (32.0 - paktmp) * 0.00695 * packf)
            s[S["COVINX"]] = covinx
            s[S["DULL"]] = np.where(nopack, 0.0, dull)
            s[S["PACKF"]] = np.where(nopack, 0.0, packf)
            s[S["PACKI"]] = np.where(nopack, 0.0, packi)
            s[S["PACKW"]] = np.where(nopack, 0.0, packw)
            s[S["PAKTMP"]] = np.where(nopack, 32.0, paktmp)
            s[S["RDENPF"]] = np.where(nopack, NAN, rdenpf)
            s[S["SKYCLR"]] = I("SKYCLR")
            s[S["XLNMLT"]] = I("XLNMLT")
            s[S["PDEPTH"]] = pdepth
            s[S["SNOCOV"]] = snocov
            s[S["NEGHTS"]] = neghts
            s[S["SNOTMP"]] = R("TSNOW")
            hr6 = calendar.hour6flag(dofirst=True)
            s[S["HR6FG"]] = 1.0 if hr6[0] > 0 else 0.0
            # OLDPREC, DEWTMP, MNEGHS, PACKWC, COMPCT, GMELTR, MOSTHT, NEGHT, RDNSN, SNOWEP, VAP, ALBEDO start at 0
        if act & B["PWATER"]:  # PWATER.py:178-260
            for nme in ("CEPS", "SURS", "UZS", "IFWS", "LZS", "GWVS"):
                s[S[nme]] = I(nme)
            agws = I("AGWS")
            s[S["AGWS"]] = np.where(agws < 0.0, 0.0, agws)
            s[S["MSUPY"]] = 0.0
            for nme in ("DEC", "SRC", "IFWK1", "IFWK2"):
                s[S[nme]] = NAN
            for nme in ("RLZRAT", "LZFRAC", "RPARM"):
                s[S[nme]] = -1.0e30
        if act & B["IWATER"]:  # IWATER.py:128-143
            s[S["RETS"]] = I("RETS")
            s[S["SURS"]] = I("SURS")
            s[S["MSUPY"]] = I("SURS")
            s[S["DEC"]] = NAN
            s[S["SRC"]] = NAN
        if act & B["SOLIDS"]:  # SOLIDS.py:64,75
            s[S["SLDS"]] = I("SLDS")
            s[S["SLD_DRYDFG"]] = 1.0
        if act & B["IWTGAS"]:  # IWTGAS.py:119-120: the raw interval-0 values, converted on the first DAYFG interval
            s[S["IWG_AWTF"]] = daily0("AWTF")
            s[S["IWG_BWTF"]] = daily0("BWTF")
        if act & B["SEDMNT"]:  # SEDMNT.py:75,80,127,145
            s[S["DETS"]] = I("DETS")
            s[S["SED_NVSI"]] = R("NVSI") * delt / 1440.0
            s[S["SED_COVER"]] = daily0("COVER")
            s[S["DRYDFG"]] = 1.0
        if act & B["PSTEMP"]:  # PSTEMP.py:78-101: deg F (default 60) to deg C; a metric run's values are deg C (default 16)
            for nme in ("SLTMP", "ULTMP", "LGTMP"):
                v = I("PST_" + nme)
                v = np.where(np.isnan(v), 16.0 if metric else 60.0, v)
                s[S[nme]] = v if metric else (v - 32.0) * 0.555
            s[S["AIRTS"]] = NAN  # set from AIRTMP[0] on the device at the first interval (PSTEMP.py:128)
        return st

    # ------------------------------------------------------------------------------------------------
    def subset(self, idx):
        """The segments `idx` (any order) as a store of their own."""
        import copy as _copy

        idx = np.asarray(idx, dtype=np.int64)
        st = _copy.copy(self)
        st.n = len(idx)
        st.par = np.ascontiguousarray(self.par[:, idx])
        st.state = np.ascontiguousarray(self.state[:, idx])
        st.flags = np.ascontiguousarray(self.flags[idx])
        st.monthly = {k: np.ascontiguousarray(v[:, idx]) for k, v in self.monthly.items()}
        return st

    def permute(self, perm):
        """Reorder the segments in place: new position i holds old segment perm[i]."""
        perm = np.asarray(perm, dtype=np.int64)
        if perm.shape != (self.n,) or not np.array_equal(np.sort(perm), np.arange(self.n)):
            raise ValueError("perm must be a permutation of range(n)")
        self.par = np.ascontiguousarray(self.par[:, perm])
        self.state = np.ascontiguousarray(self.state[:, perm])
        self.flags = np.ascontiguousarray(self.flags[perm])
        self.monthly = {k: np.ascontiguousarray(v[:, perm]) for k, v in self.monthly.items()}
        return perm

    def divergence_order(self, keys=()):
        """Divergence-aware ordering (BASELINE north_star, subsystem 3): the permutation that makes segments with equal
        option words neighbours -- a warp is 32 consecutive segments, so its lanes then take the same branches -- and,
        within equal option words, orders by the caller's `keys` (most significant first; e.g. met station, elevation
        band, temperature offset: segments that freeze, melt and get wet together)."""
        cols = [np.asarray(k) for k in reversed(tuple(keys))] + [self.flags]
        return np.lexsort(cols) if len(cols) > 1 else np.argsort(self.flags, kind="stable")

    def sort_by_flags(self, keys=()):
        """Apply divergence_order(); returns the permutation (new position i holds old segment perm[i])."""
        return self.permute(self.divergence_order(keys))


def forcing_fill(operation, act, provided):
    """Replay the wrappers' fills for absent input series in registry order.

    -> (fill {forcing name: value}, opts bits).  Raises KeyError where the reference would (an activity indexing a
    series nobody provided or filled)."""
    B = _lib.ACT_BITS
    present = set(provided)
    fill = {}
    opts = 0

    def put(names, value):
        for nme in names:
            if nme not in present:
                fill[nme] = value
                present.add(nme)

    def require(names, where):
        for nme in names:
            if nme not in present:
                raise KeyError("%s needs the time series %r" % (where, nme))

    if act & B["ATEMP"]:
        require(("GATMP", "PREC"), "ATEMP")
        present.add("AIRTMP")
    if act & B["SNOW"]:  # SNOW.py:44-46
        put(("AIRTMP", "DTMPG", "PREC", "SOLRAD", "WINMOV"), NAN)
        present.update(("RAINF", "SNOCOV", "WYIELD", "PACKI"))
    if act & B["PWATER"]:  # PWATER.py:41-48
        put(("AGWLI", "IFWLI", "LGTMP", "LZLI", "PETINP", "PREC", "SURLI", "UZLI"), 0.0)
        put(("AIRTMP", "PACKI", "PETINP", "PREC", "RAINF", "SNOCOV", "WYIELD"), NAN)
        present.update(("SURO", "SURS", "IFWO", "AGWO"))
    if act & B["IWATER"]:  # IWATER.py:35-42
        put(("AIRTMP", "PETINP", "PREC", "RAINF", "SNOCOV", "WYIELD"), NAN)
        put(("SURLI",), 0.0)
        present.update(("SURO", "SURS"))
    if act & B["SOLIDS"]:  # SOLIDS.py:24-26 (SLSLD is filled too, but never read by the loop)
        put(("SURO", "SURS", "PREC"), 0.0)
    if act & B["IWTGAS"]:  # IWTGAS.py:51-53, 92-94
        put(("AIRTMP", "WYIELD", "SURO", "SURLI"), 0.0)
        put(("SLITMP", "SLIDOX", "SLICO2"), -1.0e30)
    if act & B["SEDMNT"]:  # SEDMNT.py:86-96
        if "RAINF" not in present:
            opts |= _lib.OPT_SED_RAIN_IS_PREC
        put(("PREC", "SLSED", "SNOCOV", "SURO", "SURS"), 0.0)
    if act & B["PSTEMP"]:
        require(("AIRTMP",), "PSTEMP")
        present.update(("SLTMP", "ULTMP", "LGTMP"))
    if act & B["PWTGAS"]:  # PWTGAS.py:106-140
        put(("WYIELD", "SURO", "IFWO", "AGWO", "SURLI", "IFWLI", "AGWLI"), 0.0)
        put(("SLTMP", "ULTMP", "LGTMP", "SLITMP", "SLIDOX", "SLICO2", "ILITMP", "ILIDOX", "ILICO2", "ALITMP", "ALIDOX",
             "ALICO2"), -1.0e30)
    if act & B.get("PQUAL", 0):  # PQUAL.py:98-106 (PREC is indexed unconditionally, :125)
        require(("PREC",), "PQUAL")
        put(("SLIQSP", "ILIQC", "ALIQC", "SURO", "IFWO", "AGWO", "PERO", "WSSD", "SCRSD"), 0.0)
    if act & B.get("IQUAL", 0):  # IQUAL.py:96-98, 118-120
        require(("SURO", "SOSLD", "PREC"), "IQUAL")
        put(("SLIQSP",), -1.0e30)
    return fill, opts
