"""Synthetic replicated-segment workloads of BASELINE.json (definitions: SURVEY.md 8d "Synthetic inputs").

Base segment = test10 PERLND 1 / IMPLND 1 (hspsquared_b200/test10.py).  Segment i gets multiplicative jitter
p*(1+0.1*u_i), u_i ~ U(-1,1) from numpy.random.default_rng(seed) on the continuous parameters and the initial
storages (flags unchanged); AGWRC is clipped below 0.999, IRC to (0.3, 0.85), SHADE to [0, 0.8].  Met = test10's 1976
hourly series (hspsquared_b200/data/test10_met_1976.npz) tiled over the run with Feb-29 dropped in non-leap years,
CLOUD omitted, and per-segment forcing jitter PREC*(1+0.2u), AIRTMP+3u degF, PETINP*(1+0.1u), DTMPG = AIRTMP-5.
Everything is fp64 and deterministic in (seed, n_seg).
"""
from __future__ import annotations

import copy
import os

import numpy as np

from . import cbp, test10
from .segstore import INIT_DEFAULT, RAW_DEFAULT, SegmentStore, broadcast, effective

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "test10_met_1976.npz")
PERLND_FORCINGS = ("PREC", "PETINP", "AIRTMP", "WINMOV", "SOLRAD", "DTMPG")  # north_star list, 48 B / interval
IMPLND_FORCINGS = PERLND_FORCINGS
FULLCHAIN_FORCINGS = ("GATMP", "PREC", "PETINP", "WINMOV", "SOLRAD", "DTMPG")  # ATEMP on: gage temperature in

JITTER = {
    "PERLND": {
        "SNOW": {"PARAMETERS": ("SNOWCF", "COVIND", "MELEV", "SHADE", "CCFACT", "MWATER", "MGMELT"),
                 "STATES": ("PACKF", "PACKI", "PACKW")},
        "PWATER": {"PARAMETERS": ("LZSN", "INFILT", "LSUR", "SLSUR", "AGWRC", "UZSN", "CEPSC", "NSUR", "INTFW", "IRC",
                                  "LZETP", "DEEPFR", "AGWETP"),
                   "STATES": ("CEPS", "SURS", "UZS", "IFWS", "LZS", "AGWS", "GWVS")},
    },
    "IMPLND": {
        "SNOW": {"PARAMETERS": ("SNOWCF", "COVIND", "MELEV", "SHADE", "CCFACT", "MWATER", "MGMELT"),
                 "STATES": ("PACKF", "PACKI", "PACKW")},
        "IWATER": {"PARAMETERS": ("LSUR", "SLSUR", "NSUR", "RETSC"), "STATES": ("RETS", "SURS")},
    },
}
CLIP = {"AGWRC": (None, 0.999), "IRC": (0.3, 0.85), "SHADE": (0.0, 0.8)}


# BASELINE.json configs[3] "register / divergence stress": each option flag is drawn per segment
CBP_OPTIONS = (
    ("RTOPFG", (("PWATER", "PARAMETERS", {"RTOPFG": 1}), ("PWATER", "PARAMETERS", {"RTOPFG": 0}))),
    ("UZFG", (("PWATER", "PARAMETERS", {"UZFG": 0}), ("PWATER", "PARAMETERS", {"UZFG": 1}))),
    ("SNOPFG", (("SNOW", "FLAGS", {"SNOPFG": 0}), ("SNOW", "FLAGS", {"SNOPFG": 1}))),
    ("ICEFG", (("SNOW", "FLAGS", {"ICEFG": 1}), ("SNOW", "FLAGS", {"ICEFG": 0}))),
    ("TSOPFG", (("PSTEMP", "PARAMETERS", {"TSOPFG": 1}),
                ("PSTEMP", "PARAMETERS", {"TSOPFG": 0, "SLTVFG": 0, "ULTVFG": 0, "LGTVFG": 0, "ASLT": 14.5, "BSLT": 0.365,
                                          "ULTP1": 0.1, "ULTP2": 4.0, "LGTP1": 0.05, "LGTP2": 6.0}),
                ("PSTEMP", "PARAMETERS", {"TSOPFG": 2, "SLTVFG": 0, "ULTVFG": 0, "LGTVFG": 0, "ASLT": 14.5, "BSLT": 0.365,
                                          "ULTP1": 0.1, "ULTP2": 4.0, "LGTP1": 0.05, "LGTP2": 6.0}))),
)


def base_tables(operation, sections=None, base="test10"):
    if base == "cbp":
        if operation != "PERLND":
            raise ValueError("the CBP base segment is a PERLND")
        t = cbp.perlnd()
        t["SNOW"]["PARAMETERS"].setdefault("KMELT", 0.05)  # read when SNOPFG=1 is drawn
    else:
        t = test10.perlnd() if operation == "PERLND" else test10.implnd()
    if sections is not None:
        for a in t["ACTIVITY"]:
            t["ACTIVITY"][a] = int(a in sections)
    return t


def _clip(name, v):
    lo, hi = CLIP.get(name, (None, None))
    if lo is not None or hi is not None:
        v = np.clip(v, lo, hi)
    return v


class Ensemble:
    """n jittered replicas of a base segment; `tables_of(i)` gives segment i back as UCI tables for the oracle."""

    def __init__(self, operation, n, seed=1234, sections=None, base="test10", options=()):
        """options: names from CBP_OPTIONS whose value is drawn uniformly per segment (categorical flags)."""
        self.operation, self.n, self.seed = operation, int(n), seed
        self.base = base_tables(operation, sections, base)
        rng = np.random.default_rng(seed)
        self.options = tuple(o for o in CBP_OPTIONS if o[0] in options)
        if len(self.options) != len(tuple(options)):
            raise KeyError("unknown option in %r" % (options,))
        self.factor = {}
        for sec, groups in JITTER[operation].items():
            if not self.base["ACTIVITY"].get(sec, 0):
                continue
            for grp, names in groups.items():
                for nm in names:
                    self.factor[(sec, grp, nm)] = 1.0 + 0.1 * rng.uniform(-1.0, 1.0, self.n)
        self.u_prec = rng.uniform(-1.0, 1.0, self.n)
        self.u_airt = rng.uniform(-1.0, 1.0, self.n)
        self.u_pet = rng.uniform(-1.0, 1.0, self.n)
        # drawn last, so ensembles without options keep the numbers they always had
        self.choice = np.stack([rng.integers(0, len(alts), self.n) for _nm, alts in self.options]) if self.options \
            else np.zeros((0, self.n), dtype=np.int64)

    def _variant_tables(self, combo):
        t = copy.deepcopy(self.base)
        for (_nm, alts), k in zip(self.options, combo):
            sec, grp, kv = alts[int(k)]
            if sec in t:
                t[sec].setdefault(grp, {}).update(kv)
        return t

    def shard(self, lo, hi):
        """Segments [lo, hi) of this ensemble as an ensemble of their own (multi-GPU sharding)."""
        e = copy.copy(self)
        e.n = hi - lo
        e.factor = {k: v[lo:hi] for k, v in self.factor.items()}
        e.u_prec, e.u_airt, e.u_pet = self.u_prec[lo:hi], self.u_airt[lo:hi], self.u_pet[lo:hi]
        e.choice = self.choice[:, lo:hi]
        return e

    def take(self, idx):
        """The segments `idx` (any order) of this ensemble as an ensemble of their own."""
        idx = np.asarray(idx)
        e = copy.copy(self)
        e.n = len(idx)
        e.factor = {k: v[idx] for k, v in self.factor.items()}
        e.u_prec, e.u_airt, e.u_pet = self.u_prec[idx], self.u_airt[idx], self.u_pet[idx]
        e.choice = self.choice[:, idx]
        return e

    def tables_of(self, i):
        t = self._variant_tables(self.choice[:, i])
        for (sec, grp, nm), f in self.factor.items():
            t[sec][grp][nm] = float(_clip(nm, t[sec][grp][nm] * f[i]))
            key = "MONTHLY_" + nm
            if key in t[sec]:
                t[sec][key] = {m: float(_clip(nm, v * f[i])) for m, v in t[sec][key].items()}
        return t

    def store(self, calendar):
        """Vectorised build: effective values of the base, broadcast, jittered column-wise."""
        if self.options:
            arrs = self._variant_arrays()
        else:
            arrs = broadcast(effective(self.operation, copy.deepcopy(self.base)), self.n)
        for (sec, grp, nm), f in self.factor.items():
            base = self.base[sec][grp][nm]
            col = _clip(nm, base * f)
            if grp == "PARAMETERS":
                arrs["raw"][nm] = col
            else:
                arrs["init"][nm] = col
            if nm in arrs["mon"]:
                tab = np.array(list(self.base[sec]["MONTHLY_" + nm].values()), dtype=np.float64)
                arrs["mon"][nm] = _clip(nm, tab[:, None] * f[None, :])
        return SegmentStore.build(self.operation, arrs, calendar)

    def _variant_arrays(self):
        """broadcast() for an ensemble with categorical options: effective() once per distinct combination, then a
        gather by each segment's combination index (no per-segment Python)."""
        combos, inv = np.unique(self.choice.T, axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        effs = [effective(self.operation, self._variant_tables(c)) for c in combos]
        act = effs[0]["act"]
        if any(e["act"] != act for e in effs):
            raise ValueError("options must not change the activity mask")

        def gather(group, default, dtype):
            out = {}
            for k in set().union(*[e[group].keys() for e in effs]):
                vals = np.array([e[group].get(k, default[k] if default is not None else 0) for e in effs], dtype=dtype)
                out[k] = vals[inv]
            return out

        mon = {}
        for k in set().union(*[e["mon"].keys() for e in effs]):
            tabs = np.stack([np.asarray(e["mon"].get(k, np.zeros(12)), dtype=np.float64) for e in effs])  # [combo][12]
            mon[k] = np.ascontiguousarray(tabs[inv].T)
        return {"act": act, "n": self.n, "raw": gather("raw", RAW_DEFAULT, np.float64),
                "flag": gather("flag", None, np.uint64), "init": gather("init", INIT_DEFAULT, np.float64), "mon": mon}


_met = {}


def met_1976(operation):
    if not _met:
        z = np.load(DATA)
        _met.update({k: z[k] for k in z.files})
    pre = "P_" if operation == "PERLND" else "I_"
    return {n: _met[pre + n] for n in ("PREC", "PETINP", "AIRTMP", "WINMOV", "SOLRAD", "DTMPG")}


def met_index(calendar):
    """interval i of the run -> hour index into the (leap) 1976 series; Feb-29 dropped in non-leap years."""
    t = calendar.times[: calendar.steps]
    year = t.astype("datetime64[Y]")
    hour_of_year = ((t - year) / np.timedelta64(1, "h")).astype(np.int64)
    y = year.astype(np.int64) + 1970
    leap = ((y % 4 == 0) & (y % 100 != 0)) | (y % 400 == 0)
    feb29 = 59 * 24
    return np.where(leap | (hour_of_year < feb29), hour_of_year, hour_of_year + 24)


def shared_series(operation, calendar, t0, nsteps):
    """{name: float64[nsteps]} un-jittered series of intervals [t0, t0+nsteps) (hourly runs)."""
    if calendar.delt != 60:
        raise ValueError("synthetic met is hourly")
    idx = met_index(calendar)[t0:t0 + nsteps]
    m = met_1976(operation)
    return {k: np.ascontiguousarray(v[idx]) for k, v in m.items()}


def forcing_chunk(ens, calendar, t0, nsteps, names=PERLND_FORCINGS, out=None):
    """[nsteps][len(names)][n] float64 jittered forcings on the host."""
    s = shared_series(ens.operation, calendar, t0, nsteps)
    n = ens.n
    if out is None:
        out = np.empty((nsteps, len(names), n), dtype=np.float64)
    airt = s["AIRTMP"][:, None] + 3.0 * ens.u_airt[None, :]
    for r, nm in enumerate(names):
        if nm == "PREC":
            out[:, r, :] = s["PREC"][:, None] * (1.0 + 0.2 * ens.u_prec)[None, :]
        elif nm == "PETINP":
            out[:, r, :] = s["PETINP"][:, None] * (1.0 + 0.1 * ens.u_pet)[None, :]
        elif nm in ("AIRTMP", "GATMP"):
            out[:, r, :] = airt
        elif nm == "DTMPG":
            out[:, r, :] = airt - 5.0
        else:
            out[:, r, :] = s[nm][:, None]
    return out


def forcing_of(ens, calendar, i, t0, nsteps):
    """{name: float64[nsteps]} forcings of segment i (what forcing_chunk puts in column i), for the oracle."""
    s = shared_series(ens.operation, calendar, t0, nsteps)
    airt = s["AIRTMP"] + 3.0 * ens.u_airt[i]
    tname = "GATMP" if ens.base["ACTIVITY"].get("ATEMP", 0) else "AIRTMP"  # ATEMP on: the series is the gage's
    return {"PREC": s["PREC"] * (1.0 + 0.2 * ens.u_prec[i]), "PETINP": s["PETINP"] * (1.0 + 0.1 * ens.u_pet[i]),
            tname: airt, "DTMPG": airt - 5.0, "WINMOV": s["WINMOV"].copy(), "SOLRAD": s["SOLRAD"].copy()}
