"""Design-of-experiments runs as ONE batch: every (run, segment) pair becomes a row of the segment store.

The reference's `mainDoE.main` (src/hsp2/hsp2/mainDoE.py:19-140) re-runs the whole model once per run number, strictly one
after the other: deep copy of the UCI, `ui[table].update(run's values)` (mainDoE.py:126-131), then the usual per-segment,
per-activity dispatch.  Land segments do not interact, so for PERLND / IMPLND the runs are just more
independent segments: R runs x S segments = R*S rows advanced together by the fused kernels.

The `doe` list has the reference's format (mainDoE.py:27-68):
    [run, 'OPERATION/ACTIVITY/TABLE[/SUBTABLE]', segment, name, value]      e.g. [3, 'PERLND/PWATER/MONTHLY/CEPSC', 'P001', 'MAR', .04]
and `make_runlist` parses it exactly like mainDoE.make_runlist (mainDoE.py:202-216): the path is split at most three
times, the table name is the remainder joined with '_', values become Python floats.
"""
from __future__ import annotations

import copy
from collections import defaultdict

import numpy as np

from .calendar import Calendar
from .engine import ERRORS_OF, SECTION_OF, LandBatch, pack_forcings
from .segstore import SegmentStore
from . import _lib


def make_runlist(doe):
    """-> {run (str): {(operation, activity, segment): {table: {name: float}}}} in first-appearance order."""
    rundict = defaultdict(dict)
    for line in doe:
        run, path, segment, name, value = line[:]
        operation, module, *rest = path.split(sep="/", maxsplit=3)
        table = "_".join(rest)
        key = (operation, module, segment)
        rundict[str(run)].setdefault(key, defaultdict(dict))[table][name] = float(value)
    return rundict


def tables_for_run(operation, segment, tables, run_updates):
    """The segment's UCI tables with one run's values applied (deep copy; mainDoE.py:113,126-131)."""
    t = copy.deepcopy(tables)
    for (op, module, seg), per_table in run_updates.items():
        if op != operation or seg != segment or module not in t:
            continue
        for table, kv in per_table.items():
            t[module].setdefault(table, {}).update(kv)
    return t


def run_doe(operation, segments, forcings, siminfo, doe, save="all", device=0, chunk=None, time_unit=None):
    """Run every (run, segment) of a design of experiments for one land operation in a single batch.

    segments : {segment name: {'ACTIVITY': {...}, '<ACTIVITY>': uci[(operation, activity, segment)], ...}}
    forcings : {segment name: {series name: float64[steps]}}  (what get_timeseries builds per segment)
    Returns (results, errors): results[run][segment] = {output name: float64[steps]} for the SAVEd outputs,
    errors[run][segment] = {activity: int64 counts} -- the `errors` vectors of main.py:249-257.
    """
    if operation not in SECTION_OF:
        raise ValueError("run_doe batches PERLND or IMPLND; RCHRES runs are network-ordered and stay with the reference")
    runlist = make_runlist(doe)
    cal = Calendar(siminfo["start"], siminfo["stop"], siminfo["delt"], time_unit or siminfo.get("time_unit"))
    segs = list(segments)
    F = _lib.index("forcings")
    names = None
    rows, tabs = [], []
    for run, upd in runlist.items():
        for seg in segs:
            fn = [n for n in forcings[seg] if n in F]
            if names is None:
                names = fn
            elif set(fn) != set(names):
                raise ValueError("all segments of one DoE batch must bring the same input series")
            rows.append((run, seg))
            tabs.append(tables_for_run(operation, seg, segments[seg], upd))
    if not rows:
        return {}, {}
    store = SegmentStore.from_tables(operation, tabs, cal)
    per_seg = {seg: pack_forcings({n: np.asarray(forcings[seg][n]) for n in names}, names, 1) for seg in segs}
    forc = np.concatenate([per_seg[seg] for _run, seg in rows], axis=2)  # runs share their segment's met series
    with LandBatch(store, cal, names, save, device=device) as b:
        out = b.run(forc, chunk=chunk)
        err = b.errors()
        saved = b.save
    results, errors = defaultdict(dict), defaultdict(dict)
    for i, (run, seg) in enumerate(rows):
        results[run][seg] = {full.split("_", 1)[1]: np.ascontiguousarray(out[:, r, i]) for r, full in enumerate(saved)}
        errors[run][seg] = {a: np.array([err[k][i] for k in ERRORS_OF[a]], dtype=np.int64)
                            for a in SECTION_OF[operation] if store.act & _lib.ACT_BITS[a]}
    return dict(results), dict(errors)
