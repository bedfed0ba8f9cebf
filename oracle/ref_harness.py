"""Harness that imports the UNMODIFIED reference (respec/HSPsquared, /root/reference) in the build
container to (a) pin the C oracle and (b) generate the golden vectors committed under tests/golden/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (hspsquared_b200/), in `-m gpu` tests, in
smoke() or in bench.py imports this file: /root/reference does not exist on the GPU box.

What is patched on the harness side (never in the reference tree), and why:
  * importlib.metadata.version("hsp2") -> a string: the package is not pip-installed
    (src/hsp2/hsp2/__init__.py:11, src/hsp2/hsp2tools/__init__.py:13).
  * utilities.monthval / utilities.dayval: pandas 3 refuses `Day > Minute` (utilities.py:179,198);
    the replacements run the very same pandas pipeline minus that comparison.
"""
import importlib.metadata as _md
import os
import sys

REF_ROOT = os.environ.get("HSP2_REFERENCE", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")
# The reference as installed (unmodified, `pip install --no-deps --target baseline/_ref`): git-ignored, but it travels
# to the GPU box with the tree, where bench.py's CPU baseline times the reference's own Numba kernels (oracle/numba_baseline.py).
REF_SITE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(REF_SRC, "hsp2", "hsp2"))


def site_available():
    return os.path.isdir(os.path.join(REF_SITE, "hsp2", "hsp2"))


_loaded = {}


def load():
    """Import the reference's hot-path modules; returns a dict of module objects."""
    if _loaded:
        return _loaded
    global REF_SRC
    if not available():
        if not site_available():
            raise RuntimeError("reference tree not present at %s (nor installed under %s)" % (REF_ROOT, REF_SITE))
        REF_SRC = REF_SITE
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/hsp2_ref_numba_cache")
    _orig = _md.version

    def _version(name):
        if name == "hsp2":
            return "0.11.0a1"
        return _orig(name)

    _md.version = _version
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import hsp2.hsp2.utilities  # noqa: F401

    names = ["utilities", "ATEMP", "SNOW", "PWATER", "IWATER", "SEDMNT", "PSTEMP", "PWTGAS", "SOLIDS", "IWTGAS",
             "PQUAL", "IQUAL", "configuration"]
    import importlib

    for n in names:
        _loaded[n] = importlib.import_module("hsp2.hsp2." + n)
    _loaded["readWDM"] = sys.modules.get("hsp2.hsp2tools.readWDM") or importlib.import_module(
        "hsp2.hsp2tools.readWDM"
    )
    _patch_pandas3(_loaded)
    return _loaded


def _patch_pandas3(mods):
    """pandas-3-safe monthval/dayval: same pipeline as utilities.py:168-202, minus the offset compare."""
    import pandas as pd
    from numpy import tile
    from pandas import Series, date_range
    from pandas.tseries.offsets import Minute

    def _resample_like(ts, freq, up):
        day = pd.Timedelta(days=1)
        f = pd.Timedelta(minutes=freq.n)
        if day > f:
            return up(ts.resample(freq))
        if day < f:
            return ts.resample(freq).mean()
        return ts

    def monthval(siminfo, monthly):
        start, stop = siminfo["start"], siminfo["stop"]
        freq = Minute(siminfo["delt"])
        months = tile(monthly, stop.year - start.year + 1).astype(float)
        dr = date_range(start=f"{start.year}-01-01", end=f"{stop.year}-12-31", freq="MS")
        ts = Series(months, index=dr).resample("D").ffill()
        ts = _resample_like(ts, freq, lambda r: r.asfreq().ffill())
        return ts.truncate(start, stop).to_numpy()

    def dayval(siminfo, monthly):
        start, stop = siminfo["start"], siminfo["stop"]
        freq = Minute(siminfo["delt"])
        months = tile(monthly, stop.year - start.year + 1).astype(float)
        dr = date_range(start=f"{start.year}-01-01", end=f"{stop.year}-12-31", freq="MS")
        ts = Series(months, index=dr).resample("D").interpolate("time")
        ts = _resample_like(ts, freq, lambda r: r.ffill())
        return ts.truncate(start, stop).to_numpy()

    u = mods["utilities"]
    u.monthval = monthval
    u.dayval = dayval
    for n in ("SNOW", "PWATER", "IWATER", "SEDMNT", "PSTEMP", "PWTGAS", "ATEMP"):
        m = mods[n]
        if hasattr(m, "monthval"):
            m.monthval = monthval
        if hasattr(m, "dayval"):
            m.dayval = dayval
    # initm calls dayval through the utilities module globals, so patching utilities suffices for it.


# ------------------------------------------------------------------------------------------------
# WDM decode: the reference's pure-njit record walkers (hsp2tools/readWDM.py:166-319) driven by a
# restatement of the HDF-free part of its driver (readWDM.py:30-51, 84-140).
# ------------------------------------------------------------------------------------------------
def read_wdm(wdmfile):
    """Returns {dsn: pandas.Series(float64, DatetimeIndex with freq)} for every timeseries dataset."""
    import numpy as np
    import pandas as pd

    W = load()["readWDM"]
    iarray = np.fromfile(wdmfile, dtype=np.int32)
    farray = np.fromfile(wdmfile, dtype=np.float32)
    if iarray[0] != -998:
        raise ValueError("not a WDM file")
    nrecords = iarray[28]
    out = {}
    epoch = np.datetime64(0, "Y")
    for index in range(512, nrecords * 512, 512):
        if (iarray[index] == 0 and iarray[index + 1] == 0 and iarray[index + 2] == 0 and iarray[index + 3] == 0) or iarray[index + 5] != 1:
            continue
        dsn = int(iarray[index + 4])
        psa = iarray[index + 9]
        sacnt = iarray[index + psa - 1] if psa > 0 else 0
        pdat = iarray[index + 10]
        pdatv = iarray[index + 11]
        dattr = {"TSBDY": 1, "TSBHR": 1, "TSBMO": 1, "TSBYR": 1900, "TFILL": -999.0}
        for i in range(psa + 1, psa + 1 + 2 * sacnt, 2):
            aid = iarray[index + i]
            ptr = iarray[index + i + 1] - 1 + index
            if aid not in W.attrinfo:
                continue
            name, atype, length = W.attrinfo[aid]
            if atype == "I":
                dattr[name] = int(iarray[ptr])
            elif atype == "R":
                dattr[name] = float(farray[ptr])
        records, offsets = [], []
        for i in range(pdat + 1, pdatv - 1):
            a = iarray[index + i]
            if a != 0:
                r, o = W._splitposition(a)
                records.append(r)
                offsets.append(o)
        if not records:
            continue
        dates, values, _stop = W._process_groups(
            iarray, farray, np.asarray(records), np.asarray(offsets), dattr["TGROUP"]
        )
        dates = W._date_convert(
            np.array(dates), epoch, np.timedelta64(1, "Y"), np.timedelta64(1, "M"), np.timedelta64(1, "D"),
            np.timedelta64(1, "h"), np.timedelta64(1, "m"), np.timedelta64(1, "s"),
        )
        s = pd.Series(np.asarray(values, dtype=np.float64), index=pd.DatetimeIndex(np.array(dates)))
        try:
            s.index.freq = str(dattr["TSSTEP"]) + W.freq[dattr["TCODE"]]
        except ValueError:
            s.index.freq = None
        out[dsn] = s
    return out


def transform(series, name, how, siminfo, base=None):
    """utilities.transform (utilities.py:83-133) with the one pandas-3 incompatibility (offset / offset at
    :125 and the offset compares at :108,:116) restated for sources coarser than the run step."""
    import pandas as pd

    u = load()["utilities"]
    delt = siminfo["delt"]
    f = series.index.freq
    src = pd.Timedelta(days=f.n) if f.name == "D" else pd.Timedelta(f)
    tgt = pd.Timedelta(minutes=delt)
    if how == "DIV" and src != tgt:
        ts = series.copy()
        if ts.index[-1] < siminfo["stop"]:
            ts[siminfo["stop"]] = ts.iloc[-1]
        ts = (ts * (tgt / src)).resample(pd.tseries.offsets.Minute(delt)).ffill()
        start, stop, steps = siminfo["start"], siminfo["stop"], siminfo["steps"]
        return ts[start:stop].to_numpy().astype("float64")[0:steps]
    return (base or u.transform)(series.copy(), name, how, siminfo)


def make_siminfo(start, stop, delt=60, units=1):
    import pandas as pd
    from pandas.tseries.offsets import Minute

    start, stop = pd.Timestamp(start), pd.Timestamp(stop)
    tindex = pd.date_range(start, stop, freq=Minute(delt))[1:]  # main.py:93-95
    return {"start": start, "stop": stop, "delt": delt, "units": units, "tindex": tindex, "steps": len(tindex)}


TEST10_WDM = os.path.join(REF_ROOT, "tests", "test10", "HSP2results", "test10.wdm")

# tests/test10/HSP2results/test10.uci:891-914 (EXT SOURCES): (dsn, mfactor, tran, target member)
TEST10_EXT = {
    "PERLND": [(39, 1.0, "SAME", "PREC"), (123, 1.0, "SAME", "AIRTMP"), (41, 0.7, "DIV", "PETINP"),
               (42, 1.0, "DIV", "WINMOV"), (46, 1.0, "DIV", "SOLRAD"), (126, 1.0, "SAME", "DTMPG"),
               (135, 1.0, "SAME", "CLOUD")],
    "IMPLND": [(131, 1.0, "SAME", "PREC"), (122, 1.0, "SAME", "AIRTMP"), (41, 0.7, "DIV", "PETINP"),
               (42, 1.0, "DIV", "WINMOV"), (46, 1.0, "DIV", "SOLRAD"), (125, 1.0, "SAME", "DTMPG"),
               (135, 1.0, "SAME", "CLOUD")],
}


def test10_forcings(operation, siminfo=None):
    """Forcing arrays the way get_timeseries builds them (utilities.py:274-290): x MFACTOR, transform, sum."""
    if siminfo is None:
        siminfo = make_siminfo("1976-01-01", "1977-01-01")
    wdm = read_wdm(TEST10_WDM)
    ts = {}
    for dsn, mfactor, tran, name in TEST10_EXT[operation]:
        s = wdm[dsn].copy()
        if mfactor != 1.0:
            s = s * mfactor
            s.index.freq = wdm[dsn].index.freq
        t = transform(s, name, tran, siminfo)
        ts[name] = ts[name] + t if name in ts else t
    return ts, siminfo
