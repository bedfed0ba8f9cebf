"""Run the UNMODIFIED reference dispatcher `hsp2.hsp2.main.main()` over an in-memory IO backend so that the
drop-in activities can be exercised exactly where the reference calls them (main.py:249-250), with its own
get_timeseries in front (utilities.py:274-290) and its own save_timeseries behind (utilities.py:292-333).

TEST INFRASTRUCTURE, build container only (needs /root/reference).  Harness-side patches, none in the reference tree:
the four of oracle/ref_harness.py / SURVEY B.6 -- metadata version shim, pandas-3-safe monthval/dayval, `transform` for
daily DIV sources, and `init_state_dicts` pre-seeding state['hsp_segments'] (a model without RCHRES crashes the
reference otherwise).
"""
import collections
import copy
import os
import sys
import tempfile

from hspsquared_b200 import test10

from . import ref_harness as H

Row = collections.namedtuple("Row", "SVOLNO MFACTOR TMEMN TRAN TMEMSB")
PERLND_ACTS = ("ATEMP", "SNOW", "PWATER", "SEDMNT", "PSTEMP", "PWTGAS", "PQUAL", "MSTLAY", "PEST", "NITR", "PHOS", "TRACER")
IMPLND_ACTS = ("ATEMP", "SNOW", "IWATER", "SOLIDS", "IWTGAS", "IQUAL")


class MemoryBackend:
    """Satisfies SupportsReadUCI / ReadTS / WriteTS / WriteLogging (hsp2io/protocols.py:14-51) without HDF5."""

    def __init__(self, uci_obj, inputs):
        self.uci_obj = uci_obj
        self.inputs = inputs  # {dsn: pandas Series with index.freq}
        self.results = {}
        fd, self.file_path = tempfile.mkstemp(suffix=".h5")  # main.py:42-44 only checks that the path exists
        os.close(fd)

    def read_uci(self):
        return self.uci_obj

    def read_ts(self, category=None, operation=None, segment=None, activity=None):
        s = self.inputs[segment].copy()
        s.index.freq = self.inputs[segment].index.freq
        return s

    def write_ts(self, data_frame, category, operation=None, segment=None, activity=None, *a, **k):
        self.results[(operation, segment, activity)] = data_frame.copy()

    def write_log(self, df):
        self.log = df

    def write_versioning(self, df):
        self.versions = df

    def close(self):
        try:
            os.remove(self.file_path)
        except OSError:
            pass


CHAIN_SAVE = {
    "PSTEMP": ("AIRTC", "SLTMP", "ULTMP", "LGTMP"),
    "PWTGAS": ("SOTMP", "IOTMP", "AOTMP", "SODOX", "SOCO2", "IODOX", "IOCO2", "AODOX", "AOCO2", "SOHT", "IOHT", "AOHT", "POHT",
               "SODOXM", "SOCO2M", "IODOXM", "IOCO2M", "AODOXM", "AOCO2M", "PODOXM", "POCO2M"),
    "SEDMNT": ("DETS", "WSSD", "SCRSD", "SOSED", "COVER"),
    "SOLIDS": ("SOSLD", "SLDS"),
    "IWTGAS": ("SOTMP", "SODOX", "SOCO2", "SOHT", "SODOXM", "SOCO2M"),
}


def test10_land_model(start="1976-01-01", stop="1976-04-01", clones=1, units=1):
    """UCI object + input series for test10's PERLND P001 and IMPLND I001 (tables of hspsquared_b200/test10.py, EXT SOURCES
    of test10.uci:891-914 as in ref_harness.TEST10_EXT), SAVE tables = the SaveTable.csv defaults.
    clones > 1: P001..Pnnn and I001..Innn interleaved in the operation sequence, each with its own parameters and MFACTORs;
    every fifth PERLND runs without SNOW (another activity mask, and its PWATER finds the ICEFG the previous segment's SNOW
    left in siminfo)."""
    import pandas as pd

    m = H.load()
    UCI = sys.modules["hsp2.hsp2.uci"].UCI if "hsp2.hsp2.uci" in sys.modules else __import__("hsp2.hsp2.uci", fromlist=["UCI"]).UCI
    u = UCI()
    imp = test10.implnd()
    imp["ACTIVITY"]["IQUAL"] = 1  # test10.uci:173-177 switches IQUAL on for IMPLND 1
    imp["IQUAL"] = test10.iqual_tables()
    tables = {"PERLND": ("P001", test10.perlnd(), PERLND_ACTS), "IMPLND": ("I001", imp, IMPLND_ACTS)}
    if clones > 1:
        return _cloned_model(u, tables, clones, start, stop, units)
    rows = []
    for op, (seg, t, acts) in tables.items():
        u.uci[(op, "GENERAL", seg)] = {"ACTIVITY": {a: int(t["ACTIVITY"].get(a, 0)) for a in acts}}
        for act, tab in t.items():
            if act == "ACTIVITY":
                continue
            ui = copy.deepcopy(tab)
            ui["SAVE"] = {k: 1 for k in test10.SAVE_DEFAULT.get(act, ())}
            if units == 2 and act in CHAIN_SAVE:  # the metric run also writes the sections behind the water budget
                ui["SAVE"] = {k: 1 for k in CHAIN_SAVE[act]}
            if act == "IQUAL":  # SaveTable.csv:264-273 (all 0: written only under saveall)
                ui["SAVE"] = {k: 0 for k in ("SQO", "SOQSP", "SOQS", "SOQO", "SOQOC", "SOQUAL", "SOQC", "IQADDR", "IQADWT", "IQADEP")}
            u.uci[(op, act, seg)] = ui
        rows.append((op, seg, 60))
        u.ddext_sources[(op, seg)] = [Row(dsn, mf, name, tran, "") for dsn, mf, tran, name in H.TEST10_EXT[op]]
    u.opseq = pd.DataFrame(rows, columns=["OPERATION", "SEGMENT", "INDELT_minutes"])
    u.siminfo = {"start": pd.Timestamp(start), "stop": pd.Timestamp(stop), "units": units}  # 2: the tables read as metric values
    wdm = H.read_wdm(H.TEST10_WDM)
    return u, wdm


def _cloned_model(u, tables, clones, start, stop, units=1):
    import pandas as pd

    rows = []
    for k in range(clones):
        for op, (seg0, t0, acts) in tables.items():
            seg = "%s%03d" % (seg0[0], k + 1)
            t = copy.deepcopy(t0)
            f = 1.0 + 0.004 * k
            if op == "PERLND":
                t["ACTIVITY"].update({"PSTEMP": 0, "PWTGAS": 0})  # SNOW + PWATER (+ ATEMP off): the water budget sections
                p = t["PWATER"]["PARAMETERS"]
                p["LZSN"], p["INFILT"], p["UZSN"] = p["LZSN"] * f, p["INFILT"] / f, p["UZSN"] * f
                if k % 5 == 4:
                    t["ACTIVITY"]["SNOW"] = 0
                    p["CSNOFG"] = 0
            else:
                t["ACTIVITY"].update({"SOLIDS": 0, "IWTGAS": 0, "IQUAL": 0})
                t["IWATER"]["PARAMETERS"]["RETSC"] = t["IWATER"]["PARAMETERS"]["RETSC"] * f
                if k % 3 == 0:
                    t["SNOW"].setdefault("FLAGS", {})["ICEFG"] = 0
            t["SNOW"]["PARAMETERS"]["SNOWCF"] = t["SNOW"]["PARAMETERS"]["SNOWCF"] * f
            u.uci[(op, "GENERAL", seg)] = {"ACTIVITY": {a: int(t["ACTIVITY"].get(a, 0)) for a in acts}}
            for act, tab in t.items():
                if act == "ACTIVITY":
                    continue
                ui = copy.deepcopy(tab)
                ui["SAVE"] = {kk: 1 for kk in test10.SAVE_DEFAULT.get(act, ())}
                u.uci[(op, act, seg)] = ui
            rows.append((op, seg, 60))
            u.ddext_sources[(op, seg)] = [Row(dsn, mf * (1.0 + 0.01 * k) if name == "PREC" else mf, name, tran, "")
                                          for dsn, mf, tran, name in H.TEST10_EXT[op]]
    u.opseq = pd.DataFrame(rows, columns=["OPERATION", "SEGMENT", "INDELT_minutes"])
    u.siminfo = {"start": pd.Timestamp(start), "stop": pd.Timestamp(stop), "units": units}
    return u, H.read_wdm(H.TEST10_WDM)


def run_main(uci_obj, inputs, saveall=False):
    """-> {(operation, segment, activity): DataFrame written through IOManager.write_ts}"""
    m = H.load()
    import importlib

    main_mod = sys.modules.get("hsp2.hsp2.main") or importlib.import_module("hsp2.hsp2.main")
    io_mod = importlib.import_module("hsp2.hsp2io.io")
    util = m["utilities"]
    if not getattr(main_mod, "_hsp2g_patched", False):
        orig_init = main_mod.init_state_dicts

        def init_state_dicts():
            st = orig_init()
            st.setdefault("hsp_segments", {})
            return st

        main_mod.init_state_dicts = init_state_dicts
        orig_transform = util.transform

        def transform(ts, name, how, siminfo):
            return H.transform(ts, name, how, siminfo, base=orig_transform)

        util.transform = transform
        main_mod._hsp2g_patched = True
    backend = MemoryBackend(copy.deepcopy(uci_obj), inputs)
    try:
        main_mod.main(io_mod.IOManager(backend), saveall=saveall, jupyterlab=False)
    finally:
        backend.close()
    return backend.results
