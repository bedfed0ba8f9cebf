#!/usr/bin/env python
"""The reference's OWN Numba implementation of the headline path, timed on this host's cores: serial and multiprocess.

TEST / MEASUREMENT INFRASTRUCTURE (only bench.py's cpu_baseline leg runs it).  It imports the UNMODIFIED reference --
/root/reference in the build container, else the copy installed under baseline/_ref (oracle/ref_harness.py) -- and runs
the reference wrappers `snow()` and `pwater()` (SNOW.py:27, PWATER.py:27) for segments of bench.py's synthetic ensemble
(hspsquared_b200/synth.py: test10 PERLND 1 with jittered parameters and forcings), exactly as `hsp2.main` calls them
(main.py:249-250).  Two figures per run: wrapper-inclusive (what a user of the reference pays per segment: pandas calendar
glue + typed-Dict construction + the kernels) and kernel-only (the time spent inside the @njit cores `_snow_` / `_pwater_`,
measured by timing those calls inside the wrappers), both in segment-timesteps per second.  The multiprocess figure runs the
same loop in a multiprocessing.Pool over segment blocks (the kernels are compiled without nogil, so threads do not help:
SURVEY.md 8d).  JIT compilation / cache loading happens in an untimed warm-up call per process.

    python -m oracle.numba_baseline --procs 16 --segments-per-proc 24 --steps 8784      -> one JSON line
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_state = {}


def _init():
    """per process: import the reference, wrap its cores with timers, build the ensemble"""
    if _state:
        return
    import numpy as np
    import pandas as pd

    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import Calendar
    from oracle import ref_harness as H

    m = H.load()
    acc = {"kernel_s": 0.0}
    for mod, name in ((m["SNOW"], "_snow_"), (m["PWATER"], "_pwater_")):
        orig = getattr(mod, name)

        def timed(ui, ts, _orig=orig):
            t0 = time.perf_counter()
            r = _orig(ui, ts)
            acc["kernel_s"] += time.perf_counter() - t0
            return r

        setattr(mod, name, timed)
    _state.update(np=np, pd=pd, synth=synth, Calendar=Calendar, snow=m["SNOW"].snow, pwater=m["PWATER"].pwater, acc=acc,
                  where=H.REF_SRC)


def _run_block(args):
    """segments [lo, hi) of the ensemble over `steps` hourly intervals from 1976-01-01 -> (kernel s, wall s, n, checksum)"""
    lo, hi, steps, nseg_total, seed = args
    _init()
    from numba import types
    from numba.typed import Dict

    from oracle import cases as CASES

    np, pd, synth = _state["np"], _state["pd"], _state["synth"]
    start = np.datetime64("1976-01-01T00:00")
    cal = _state["Calendar"](start, start + np.timedelta64(steps, "h"), 60, "ns")
    ens = _state.get(("ens", nseg_total, seed))
    if ens is None:
        ens = _state[("ens", nseg_total, seed)] = synth.Ensemble("PERLND", nseg_total, seed=seed, sections=("SNOW", "PWATER"))
    acts = {"SNOW": _state["snow"], "PWATER": _state["pwater"]}
    _state["acc"]["kernel_s"] = 0.0
    chk, wrap_s = 0.0, 0.0
    for i in range(lo, hi):
        si = {"start": pd.Timestamp(str(cal.start)), "stop": pd.Timestamp(str(cal.start + np.timedelta64(steps, "h"))), "delt": 60,
              "units": 1, "steps": steps}
        ts = Dict.empty(key_type=types.unicode_type, value_type=types.float64[:])
        for k, v in synth.forcing_of(ens, cal, i, 0, steps).items():
            ts[k] = v
        tables = ens.tables_of(i)
        t0 = time.perf_counter()  # the reference's work only: snow() then pwater() of one segment (inputs were built outside)
        CASES.run_segment("PERLND", tables, ts, si, acts)
        wrap_s += time.perf_counter() - t0
        chk += float(ts["PERO"].sum())
    return _state["acc"]["kernel_s"], wrap_s, hi - lo, chk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--segments-per-proc", type=int, default=16)
    ap.add_argument("--serial-segments", type=int, default=24)
    ap.add_argument("--steps", type=int, default=8784, help="intervals per segment (the reference runs a segment's whole series)")
    ap.add_argument("--repeats", type=int, default=1, help="timed passes of the multiprocess job (bench.py --steps)")
    ap.add_argument("--warm", type=int, default=1, help="untimed passes before them (bench.py --warmup)")
    a = ap.parse_args()
    import multiprocessing as mp

    from oracle import ref_harness as H

    if not (H.available() or H.site_available()):
        print(json.dumps({"unavailable": "the reference is neither at /root/reference nor installed under baseline/_ref"}))
        return
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/hsp2_ref_numba_cache")
    seed, steps = 1234, a.steps
    # ---- serial: this process
    serial = None
    if a.serial_segments > 0:
        _run_block((0, 2, min(steps, 240), 64, seed))  # warm-up: JIT / cache load
        ks, ws, n, chk = _run_block((0, a.serial_segments, steps, max(64, a.serial_segments), seed))
        serial = {"segments": n, "intervals": steps, "kernel_only": n * steps / ks, "wrapper_inclusive": n * steps / ws,
                  "kernel_s": round(ks, 3), "wrapper_s": round(ws, 3)}
    # ---- multiprocess over segment blocks: all P processes work at the same time on equal blocks, a pass takes as long as the
    # slowest of them
    P = max(1, a.procs)
    ntot = P * a.segments_per_proc
    ctx = mp.get_context("spawn")
    passes = []
    with ctx.Pool(P) as pool:
        pool.map(_run_block, [(k, k + 1, min(steps, 240), ntot, seed) for k in range(P)], chunksize=1)  # JIT / cache load per worker
        for it in range(a.warm + a.repeats):
            t0 = time.perf_counter()
            res = pool.map(_run_block, [(k * a.segments_per_proc, (k + 1) * a.segments_per_proc, steps, ntot, seed)
                                        for k in range(P)], chunksize=1)
            wall = time.perf_counter() - t0
            if it >= a.warm:
                passes.append({"kernel_s": max(r[0] for r in res), "wrapper_s": max(r[1] for r in res), "wall_s": wall,
                               "checksum": sum(r[3] for r in res)})
    work = ntot * steps * len(passes)
    multi = {"processes": P, "segments": ntot, "intervals": steps, "passes": len(passes),
             "kernel_only": work / sum(p["kernel_s"] for p in passes),
             "wrapper_inclusive": work / sum(p["wrapper_s"] for p in passes),
             "kernel_s_per_pass": round(sum(p["kernel_s"] for p in passes) / len(passes), 4),
             "wrapper_s_per_pass": round(sum(p["wrapper_s"] for p in passes) / len(passes), 4),
             "wall_s_per_pass": round(sum(p["wall_s"] for p in passes) / len(passes), 4), "checksum": passes[-1]["checksum"]}
    import numba

    print(json.dumps({"kind": "reference", "implementation": "respec/HSPsquared v0.11.0a1, unmodified: snow() and pwater() "
                      "(SNOW.py:27, PWATER.py:27) with their @njit cores _snow_ / _pwater_",
                      "reference_loaded_from": _state.get("where") or H.REF_SRC, "numba": numba.__version__,
                      "unit": "segment-timesteps/s", "serial": serial, "multiprocess": multi}))


if __name__ == "__main__":
    main()
