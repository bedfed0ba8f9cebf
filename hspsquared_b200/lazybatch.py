"""Lazy batching behind the reference's per-(segment, activity) dispatch (SURVEY.md 8b "granularity mismatch", 7 step 6).

`hsp2.main` walks the operation sequence one segment at a time and calls one activity function at a time
(main.py:91-275, call at :249-250).  A GPU wants all segments at once.  This module reconciles the two without touching a
reference file: the FIRST land-activity call of a segment that is not yet in the cache looks at the dispatcher that is
calling it (the frame of `main()`: its `uci` mapping, `opseq`, `ddext_sources`, `io_manager`), takes the next window of
segments of the same operation from the operation sequence, builds their forcings with the reference's own
`get_timeseries` (utilities.py:274-290, the very function object `main` uses), runs ALL their active sections fused in
ONE batch per chunk (engine.LandBatch), and keeps the outputs.  Every later `(segment, activity)` call of the window is
served from those columns: it writes the section's output series into the caller's `ts`, returns the section's error
vector, and the wrapper around it keeps the reference's side effects (helper series, siminfo['ICEFG'], CSNOFG).

Safety: a call is served from the cache only if the `ui` it received IS the table object of the cached segment and every
input series in its `ts` equals the series the batch was fed; anything else (lateral inflows linked in from another
operation, GENER inputs, a caller that is not `hsp2.main`) falls back to the one-segment path of activities.py.

The window is bounded by a memory budget (all outputs of the window's segments stay on the host until they have been
served): LAZY["max_bytes"], LAZY["max_segments"].  LAZY["enabled"] = False switches the module off.
"""
from __future__ import annotations

import sys

import numpy as np

from . import _lib
from .engine import ERRORS_OF, SECTION_OF, LandBatch, output_names
from .segstore import SegmentStore, effective, forcing_fill, stack

LAZY = {"enabled": True, "max_bytes": 3 << 30, "max_segments": 65536, "min_segments": 2}
STATS = {"windows": 0, "segments": 0, "served": 0, "fallback": 0, "launches": 0}
LAND_SECTIONS = {"PERLND": ("ATEMP", "SNOW", "PWATER", "SEDMNT", "PSTEMP", "PWTGAS"),
                 "IMPLND": ("ATEMP", "SNOW", "IWATER", "SOLIDS", "IWTGAS")}
_NEED = ("uci", "opseq", "ddext_sources", "io_manager", "operation", "segment")
_caches = {}


def main_context(max_depth=8):
    """The local variables of the dispatcher (`hsp2.hsp2.main.main`, or anything shaped like it) that is calling the current
    activity function, or None."""
    f = sys._getframe(2)
    for _ in range(max_depth):
        if f is None:
            return None
        loc = f.f_locals
        if all(k in loc for k in _NEED):
            return loc, f.f_globals
        f = f.f_back
    return None


class _Window:
    def __init__(self):
        self.seg = {}  # segment -> dict(ui={section: table object}, inputs={name: array}, col=i, group=g)
        self.groups = []  # dict(save=[...], out=[steps][ns][n], errors={id: int64[n]})


class _RunCache:
    def __init__(self, uci):
        self.uci = uci  # keeps the mapping alive, so its id() cannot be reused while the cache exists
        self.windows = {}  # operation -> _Window


def _cache_for(uci):
    c = _caches.get(id(uci))
    if c is None or c.uci is not uci:
        _caches.clear()  # one run at a time: a new uci mapping means a new main() call
        c = _caches[id(uci)] = _RunCache(uci)
    return c


def serve(section, operation, siminfo, ui, ts, calendar, device=0):
    """-> error vector of `section` for the calling segment with the section's outputs written into ts, or None when this call
    cannot be served from a batch (the caller then runs the one-segment path)."""
    if not LAZY["enabled"]:
        return None
    ctx = main_context()
    if ctx is None:
        return None
    loc, glob = ctx
    uci, seg = loc["uci"], loc["segment"]
    operation = loc["operation"]  # atemp / snow serve both land operations: the dispatcher knows which one this is
    if section not in LAND_SECTIONS.get(operation, ()) or uci.get((operation, section, seg)) is not ui:
        return None
    cache = _cache_for(uci)
    w = cache.windows.get(operation)
    if w is None or seg not in w.seg:
        try:
            w = _build_window(loc, glob, operation, seg, siminfo, calendar, device)
        except _NotBatchable:
            w = None
        cache.windows[operation] = w
        if w is None or seg not in w.seg:
            STATS["fallback"] += 1
            return None
    e = w.seg[seg]
    if section not in e["sections"] or e["ui"].get(section) is not ui:
        STATS["fallback"] += 1
        return None
    # every input of this section must be the series the batch was fed (main builds ts afresh per segment, then links in
    # flows from other operations: anything the batch did not see disqualifies the cached result)
    from .activities import INPUTS

    g = w.groups[e["group"]]
    i = e["col"]
    for name in INPUTS[section]:
        if name in e["written"]:
            continue
        if name in e["inputs"]:
            ok = name in ts and np.array_equal(np.asarray(ts[name]), e["inputs"][name], equal_nan=True)
        elif name in ts:
            # a series the batch was not given: acceptable only if it is the constant an earlier wrapper of this segment
            # planted for an absent input (PWATER.py:41-48 and friends) -- the very value the fused kernels read for it
            v, fill = np.asarray(ts[name]), g["fills"].get(name)
            ok = fill is not None and bool(np.all(np.isnan(v)) if np.isnan(fill) else np.all(v == fill))
        else:
            ok = True
        if not ok:
            STATS["fallback"] += 1
            return None
    for r, full in enumerate(g["save"]):
        sec, name = full.split("_", 1)
        if sec == section:
            ts[name] = np.ascontiguousarray(g["out"][:, r, i])
            e["written"].add(name)
    STATS["served"] += 1
    return np.array([g["errors"][k][i] for k in ERRORS_OF[section]], dtype=np.int64)


class _NotBatchable(Exception):
    pass


def _active_sections(operation, flags):
    # main.py:123-128: registry order, skipped when the flag is present and falsy
    return tuple(a for a in LAND_SECTIONS[operation] if (a not in flags) or flags[a]) if flags is not None else ()


def _build_window(loc, glob, operation, first_seg, siminfo, calendar, device):
    uci, opseq, ddext, io_manager = loc["uci"], loc["opseq"], loc["ddext_sources"], loc["io_manager"]
    ddlinks = loc.get("ddlinks") or {}
    get_timeseries = glob.get("get_timeseries")
    if get_timeseries is None:
        raise _NotBatchable()
    rows = [(op, seg, delt) for _i, op, seg, delt in opseq.itertuples()]
    try:
        start = next(k for k, (op, seg, _d) in enumerate(rows) if op == operation and seg == first_seg)
    except StopIteration:
        raise _NotBatchable()
    delt = rows[start][2]
    steps = int(siminfo["steps"])
    F = _lib.index("forcings")
    # siminfo['ICEFG'] is written by every snow() call and read by the next pwater() (SNOW.py:60-63, PWATER.py:87), across
    # segments and operations: replay that sequence over the window
    icefg_known = "ICEFG" in siminfo
    icefg = siminfo.get("ICEFG", 0)
    picked, budget = [], LAZY["max_bytes"]
    for op, seg, d in rows[start:]:
        gen = uci.get((op, "GENERAL", seg))
        flags = gen.get("ACTIVITY") if gen else None
        snow_on = op in LAND_SECTIONS and flags is not None and flags.get("SNOW", 0) and (op, "SNOW", seg) in uci
        if op == operation and d == delt:
            secs = tuple(a for a in _active_sections(op, flags) if (op, a, seg) in uci)
            if secs and not ddlinks.get(seg):  # a segment fed by another operation's flows runs one at a time
                n_out = len(output_names(op, sum(_lib.ACT_BITS[a] for a in secs)))
                cost = steps * 8 * (n_out + 8)
                if picked and (budget < cost or len(picked) >= LAZY["max_segments"]):
                    break
                budget -= cost
                # what pwater() of this segment will find in siminfo: its own snow() ran just before it
                seen = (gen_icefg(uci, op, seg) if snow_on else icefg) if (icefg_known or snow_on) else 0.0
                picked.append((seg, secs, seen))
        if snow_on:
            icefg, icefg_known = gen_icefg(uci, op, seg), True
    if len(picked) < LAZY["min_segments"]:
        raise _NotBatchable()
    w = _Window()
    groups = {}
    for seg, secs, seen in picked:
        ts = get_timeseries(io_manager, ddext[(operation, seg)], siminfo)
        names = tuple(sorted((n for n in ts.keys() if n in F), key=lambda n: F[n]))
        tables = {"ACTIVITY": {a: 1 for a in secs}}
        for a in secs:
            tables[a] = uci[(operation, a, seg)]
        try:
            eff = effective(operation, tables, pw_icefg=seen, ext_names=[k for k in ts.keys()])
        except (NotImplementedError, KeyError):
            continue  # the one-segment path raises the same error at the call that owns it
        inputs = {n: np.array(ts[n], dtype=np.float64, copy=True) for n in names}
        if any(len(v) != steps for v in inputs.values()):
            continue
        groups.setdefault((eff["act"], names), []).append((seg, secs, eff, inputs))
    for (act, names), members in groups.items():
        if len(members) < 1:
            continue
        store = SegmentStore.build(operation, stack([m[2] for m in members]), calendar, units=int(siminfo.get("units", 1)))
        save = output_names(operation, act)
        forcing = np.empty((steps, len(names), len(members)))
        for i, m in enumerate(members):
            for r, nme in enumerate(names):
                forcing[:, r, i] = m[3][nme]
        with LandBatch(store, calendar, list(names), save, device=device, order=None) as batch:
            out = batch.run(forcing, chunk=max(1, min(steps, (256 << 20) // max(1, 8 * len(members) * (len(names) + len(save))))))
            errs = batch.errors()
            launches, _ms = batch.kernel_stats()
        STATS["launches"] += int(launches)
        gi = len(w.groups)
        fills, _opts = forcing_fill(operation, act, list(names))
        w.groups.append({"save": list(save), "out": out, "errors": errs, "fills": fills})
        for i, (seg, secs, _eff, inputs) in enumerate(members):
            w.seg[seg] = {"sections": set(secs), "ui": {a: uci[(operation, a, seg)] for a in secs}, "inputs": inputs, "col": i,
                          "group": gi, "written": set()}
    STATS["windows"] += 1
    STATS["segments"] += len(w.seg)
    return w


def gen_icefg(uci, op, seg):
    u = uci[(op, "SNOW", seg)]
    return u["FLAGS"]["ICEFG"] if "FLAGS" in u and "ICEFG" in u["FLAGS"] else 0


def reset():
    _caches.clear()
    for k in STATS:
        STATS[k] = 0
