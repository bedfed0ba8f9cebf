"""ctypes binding of the CPU oracle (oracle/hsp2_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The functions mirror the shape of the reference's Numba cores -- `_snow_(ui, ts) -> errors` with the
outputs written into `ts` (SNOW.py:91-92, PWATER.py:102-103, IWATER.py:75-76, ...) -- so that parity tests
read like the reference's own tests.  `ui` is a plain {name: number} dict, `ts` a {name: float64 array} dict.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhsp2oracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement (gcc only; seconds)."""
    src = [os.path.join(_HERE, f) for f in ("hsp2_oracle.c", "hsp2_oracle.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.hsp2o_fields.restype = C.c_char_p
        _lib.hsp2o_fields.argtypes = [C.c_char_p]
    return _lib


def fields(record):
    s = lib().hsp2o_fields(record.encode()).decode()
    return [x for x in s.split(",") if x]


_dp = C.POINTER(C.c_double)


def _records(name):
    ui_f, in_f, out_f = fields(name + "_ui"), fields(name + "_in"), fields(name + "_out")
    UI = type(name + "_ui", (C.Structure,), {"_fields_": [(f, C.c_double) for f in ui_f]})
    IN = type(name + "_in", (C.Structure,), {"_fields_": [(f, _dp) for f in in_f]})
    OUT = type(name + "_out", (C.Structure,), {"_fields_": [(f, _dp) for f in out_f]})
    return UI, IN, OUT, ui_f, in_f, out_f


_cache = {}


def _run(name, ui, ts, nsteps, nerr, in_alias=None, ui_defaults=None):
    """Generic driver: fill the three records from dicts, call hsp2o_<name>, put outputs in ts."""
    if name not in _cache:
        _cache[name] = _records(name)
    UI, IN, OUT, ui_f, in_f, out_f = _cache[name]
    u = UI()
    for f in ui_f:
        if f in ui:
            setattr(u, f, float(ui[f]))
        elif ui_defaults and f in ui_defaults:
            setattr(u, f, float(ui_defaults[f]))
        else:
            raise KeyError("oracle %s: ui[%r] missing" % (name, f))
    keep = []
    i = IN()
    for f in in_f:
        key = in_alias.get(f, f) if in_alias else f
        a = np.ascontiguousarray(ts[key], dtype=np.float64)
        if a.shape[0] < nsteps:
            raise ValueError("oracle %s: ts[%r] shorter than nsteps" % (name, key))
        keep.append(a)
        setattr(i, f, a.ctypes.data_as(_dp))
    o = OUT()
    outs = {}
    for f in out_f:
        a = np.empty(nsteps, dtype=np.float64)
        outs[f] = a
        setattr(o, f, a.ctypes.data_as(_dp))
    fn = getattr(lib(), "hsp2o_" + name)
    if nerr is None:
        fn(C.byref(u), C.byref(i), C.byref(o), C.c_int64(nsteps))
        errors = np.zeros(0, dtype=np.int64)
    else:
        errors = np.zeros(nerr, dtype=np.int64)
        fn(C.byref(u), C.byref(i), C.byref(o), C.c_int64(nsteps), errors.ctypes.data_as(C.POINTER(C.c_int64)))
    for f, a in outs.items():
        ts[f] = a
    return errors


def _atemp_(ui, ts):
    return _run("atemp", ui, ts, int(ui["steps"]), None)


def _snow_(ui, ts):
    return _run("snow", ui, ts, int(ui["steps"]), 1, ui_defaults={"ICEFG": 0, "SNOPFG": 0, "uunits": 1})


def _pwater_(ui, ts):
    if int(ui.get("HWTFG", 0)):
        raise NotImplementedError("HWTFG=1 is not implemented in the reference (PWATER.py:107-110)")
    d = {"ICEFG": 0, "FZG": 1.0, "FZGL": 0.1, "INFILT_MFAC": 1.0}
    if "CSNOFG" not in ui:  # PWATER.py:159-176 defaults when the PWAT-PARM1 flags are absent
        d.update({"CSNOFG": 0, "IFFCFG": 1, "IFRDFG": 0, "RTOPFG": 0, "UZFG": 0, "VLEFG": 0})
    return _run("pwater", ui, ts, int(ui["steps"]), 10, ui_defaults=d)


def _iwater_(ui, ts):
    return _run("iwater", ui, ts, int(ui["steps"]), 1, ui_defaults={"RTLIFG": 0, "CSNOFG": 0, "RTOPFG": 0})


def _sedmnt_(ui, ts):
    ts = ts
    if "SLSED" not in ts:  # SEDMNT.py:94-96 zero fills
        ts["SLSED"] = np.zeros(int(ui["simlen"]))
    for n in ("PREC", "SNOCOV", "SURO", "SURS"):
        if n not in ts:
            ts[n] = np.zeros(int(ui["simlen"]))
    ts["_RAIN"] = ts["RAINF"] if "RAINF" in ts else ts["PREC"]  # SEDMNT.py:86-89
    e = _run("sedmnt", ui, ts, int(ui["simlen"]), None, in_alias={"RAIN": "_RAIN"},
             ui_defaults={"VSIVFG": 0, "SDOPFG": 0, "uunits": 1})
    del ts["_RAIN"]
    return e


def _pstemp_(ui, ts):
    t0 = 16.0 if ui.get("uunits", 1) == 2 else 60.0  # PSTEMP.py:80-88
    return _run("pstemp", ui, ts, int(ui["simlen"]), 6,
                ui_defaults={"TSOPFG": 0, "AIRTC": t0, "SLTMP": t0, "ULTMP": t0, "LGTMP": t0, "uunits": 1})


def _pwtgas_(ui, ts):
    n = int(ui["simlen"])
    for name in ("WYIELD", "SURO", "IFWO", "AGWO", "SURLI", "IFWLI", "AGWLI"):  # PWTGAS.py:106-108
        if name not in ts:
            ts[name] = np.zeros(n)
    for name in ("SLTMP", "ULTMP", "LGTMP", "SLITMP", "SLIDOX", "SLICO2", "ILITMP", "ILIDOX", "ILICO2",
                 "ALITMP", "ALIDOX", "ALICO2"):  # PWTGAS.py:117-140
        if name not in ts:
            ts[name] = np.full(n, -1.0e30)
    return _run("pwtgas", ui, ts, n, None, ui_defaults={"GDVFG": 0, "uunits": 1})


def _solids_(ui, ts):
    n = int(ui["simlen"])
    for nm in ("SURO", "SURS", "PREC", "SLSLD"):  # SOLIDS.py:24-26 zero fills
        if nm not in ts:
            ts[nm] = np.zeros(n)
    return _run("solids", ui, ts, n, None, ui_defaults={"SDOPFG": 0})


def _iwtgas_(ui, ts):
    n = int(ui["simlen"])
    for nm in ("SLITMP", "SLIDOX", "SLICO2"):  # IWTGAS.py:92-94
        if nm not in ts:
            ts[nm] = np.full(n, -1.0e30)
    return _run("iwtgas", ui, ts, n, None, ui_defaults={"uunits": 1})


def _pqual_(ui, ts):
    """_pqual_(ui, ts) (PQUAL.py:147-397): constituents 1..nquals with per-constituent ui / ts keys; outputs
    ts['PQUAL<k>_<NAME>']; errors[0] counts constituents that take atmospheric deposition without QSOFG (:171-174)."""
    n, nq = int(ui["simlen"]), int(ui["nquals"])
    errors = np.zeros(int(ui["errlen"]), dtype=np.int64)
    for k in range(1, nq + 1):
        K = str(k)
        if ui["QSOFG" + K] == 0 and (ui["pqadfgf" + K] != 0 or ui["pqadfgc" + K] != 0):
            errors[0] += 1
        u = {"delt60": ui["delt60"], "SLIFAC": ui["SLIFAC"], "ILIFAC": ui["ILIFAC"], "ALIFAC": ui["ALIFAC"]}
        for nm in ("QSDFG", "QSOFG", "QIFWFG", "QAGWFG", "VIQCFG", "VAQCFG", "SQO", "WSQOP"):
            u[nm] = ui[nm + K]
        t = {nm: ts[nm] for nm in ("SURO", "IFWO", "AGWO", "PERO", "WSSD", "SCRSD", "PREC", "SLIQSP", "ILIQC", "ALIQC", "DAYFG")}
        for nm in ("POTFW", "POTFS", "ACQOP", "REMQOP", "IOQCP", "AOQCP", "PQADFX", "PQADCN"):
            t[nm] = ts[nm + K]
        _run("pqual", u, t, n, None)
        for f in fields("pqual_out"):
            ts["PQUAL%s_%s" % (K, f)] = t[f]
    return errors


def _iqual_(ui, ts):
    """_iqual_(ui, ts) (IQUAL.py:109-284); outputs ts['IQUAL<k>_<NAME>']."""
    n, nq = int(ui["simlen"]), int(ui["nquals"])
    errors = np.zeros(int(ui["errlen"]), dtype=np.int64)
    for k in range(1, nq + 1):
        K = str(k)
        if ui["QSOFG" + K] == 0 and (ui["iqadfgf" + K] != 0 or ui["iqadfgc" + K] != 0):
            errors[0] += 1
        u = {"delt60": ui["delt60"], "SLIFAC": ui["SLIFAC"], "QSDFG": ui["QSDFG" + K], "QSOFG": ui["QSOFG" + K],
             "SQO": ui["sqo" + K], "WSQOP": ui["wsqop" + K]}
        t = {nm: ts[nm] for nm in ("SURO", "SOSLD", "PREC", "SLIQSP", "DAYFG")}
        for nm in ("POTFW", "ACQOP", "REMQOP", "IQADFX", "IQADCN"):
            t[nm] = ts[nm + K]
        _run("iqual", u, t, n, None)
        for f in fields("iqual_out"):
            ts["IQUAL%s_%s" % (K, f)] = t[f]
    return errors


def perlnd_batch(nsteps, delt, forc, par, mon, cal, nthreads=0):
    """SNOW->PWATER over independent segments with OpenMP; returns (sums[nseg,24], errors[nseg,11], threads)."""
    forc = np.ascontiguousarray(forc, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    mon = np.ascontiguousarray(mon, dtype=np.float64)
    cal = np.ascontiguousarray(cal, dtype=np.float64)
    nseg = par.shape[0]
    assert forc.shape == (nseg, 6, nsteps) and mon.shape == (nseg, 4, nsteps + 1) and cal.shape == (4, nsteps + 1)
    assert par.shape[1] == len(fields("batch_par"))
    sums = np.zeros((nseg, 24))
    errors = np.zeros((nseg, 11), dtype=np.int64)
    L = lib()
    L.hsp2o_perlnd_batch.restype = C.c_int
    used = L.hsp2o_perlnd_batch(
        C.c_int64(nseg), C.c_int64(nsteps), C.c_double(delt), forc.ctypes.data_as(_dp), par.ctypes.data_as(_dp),
        mon.ctypes.data_as(_dp), cal.ctypes.data_as(_dp), sums.ctypes.data_as(_dp),
        errors.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(nthreads))
    return sums, errors, used
