"""Shared parity harness: run a case of oracle/cases.py through (a) the CPU oracle, (b) the fused LandBatch path,
(c) the drop-in activities, and compare.  Used by the CPU suite (against the host-compiled test double, bit-exact) and
by the `-m gpu` suite (against the CUDA library through the C ABI, <= 1e-9 relative, integers exact)."""
import copy
import os
import warnings

import numpy as np

from hspsquared_b200 import _lib, activities, synth, test10
from hspsquared_b200.calendar import Calendar
from hspsquared_b200.engine import ERRORS_OF, LandBatch, pack_forcings
from hspsquared_b200.ingest import MetStore, SourceRow, resample
from hspsquared_b200.wdm import WdmSeries
from hspsquared_b200.segstore import SegmentStore
from oracle import cases as CASES
from oracle import wrappers as W

ORACLE_ACTS = {"ATEMP": W.atemp, "SNOW": W.snow, "PWATER": W.pwater, "IWATER": W.iwater, "SEDMNT": W.sedmnt,
               "PSTEMP": W.pstemp, "PWTGAS": W.pwtgas, "SOLIDS": W.solids, "IWTGAS": W.iwtgas, "PQUAL": W.pqual, "IQUAL": W.iqual}
RTOL = 1e-9  # BASELINE.json north_star: max relative error <= 1e-9 on all saved flows and storages


def host_libm_is_fma_build():
    """glibc picks its FMA builds of pow / exp / log by ifunc when the CPU has FMA and AVX2 (sysdeps/x86_64/fpu/multiarch);
    the device math restates THOSE (hspsquared_b200/csrc/glibc_math_core.h).  On such a host -- every box of this pool --
    the oracle and the GPU agree bit for bit; elsewhere the comparison falls back to RTOL."""
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((line for line in f if line.startswith("flags")), "")
        return " fma " in flags + " " and " avx2 " in flags + " "
    except OSError:
        return False


GPU_EXACT = host_libm_is_fma_build()


def run_oracle(case, time_unit="us"):
    si = CASES.siminfo_for(case)
    si["time_unit"] = time_unit
    ts = dict(CASES.forcings(case))
    errs = CASES.run_segment(case["operation"], copy.deepcopy(case["tables"]), ts, si, ORACLE_ACTS)
    return ts, errs


def run_fused(case, time_unit="us", chunk=None, replicate=1, device=0, sub_steps=None, save="all", max_steps=None,
              out_dtype=None, jit=None):
    """All active sections fused in one kernel; `replicate` copies of the segment (tests padding / multi-block).  A PQUAL /
    IQUAL section runs as a batch of (segment, constituent) rows fed by the land batch's outputs (engine.QualBatch)."""
    si = CASES.siminfo_for(case)
    cal = Calendar(si["start"], si["stop"], si["delt"], time_unit)
    forc = CASES.forcings(case)
    if max_steps is not None:
        forc = {k: v[:max_steps] for k, v in forc.items()}
    F = _lib.index("forcings")
    op = case["operation"]
    qsec = "PQUAL" if op == "PERLND" else "IQUAL"
    has_qual = bool(case["tables"]["ACTIVITY"].get(qsec, 0))
    land_on = any(v for k, v in case["tables"]["ACTIVITY"].items() if k != qsec)
    ts, errs, kname = {}, {}, None
    if land_on:
        names = [n for n in forc if n in F and n not in ("SLIQSP", "ILIQC", "ALIQC")]
        tables = [copy.deepcopy(case["tables"]) for _ in range(replicate)]
        store = SegmentStore.from_tables(op, tables, cal, units=si.get("units", 1))
        with LandBatch(store, cal, names, save, device=device, jit=jit) as b:
            if sub_steps is not None:
                b.set_schedule(sub_steps)
            if out_dtype is not None:
                b.set_output_dtype(out_dtype)
            out = b.run(pack_forcings(forc, names, replicate), chunk=chunk)
            e = b.errors()
            kname = b.kernel_name()
            saved = b.save
        for a in CASES.ORDER[op]:
            if a != qsec and store.act & _lib.ACT_BITS[a]:
                errs[a] = np.stack([e[k] for k in ERRORS_OF[a]]) if ERRORS_OF[a] else np.zeros((0, replicate), dtype=np.int64)
        ts = {full.split("_", 1)[1]: out[:, r, :] for r, full in enumerate(saved)}
    if has_qual and save == "all" and out_dtype is None:
        from hspsquared_b200.engine import QualBatch
        from hspsquared_b200.qualstore import IQUAL_INPUTS, PQUAL_INPUTS

        series = dict(forc)
        series.update(ts)  # the land sections' outputs override nothing the constituents read from outside
        qn = [n for n in (PQUAL_INPUTS if op == "PERLND" else IQUAL_INPUTS) if n in series]
        ucis = [copy.deepcopy(case["tables"][qsec]) for _ in range(replicate)]
        with QualBatch(op, ucis, cal, qn, "all", device=device) as q:
            if sub_steps is not None:
                q.set_schedule(sub_steps)
            dep = {(i, key[:-2]): v for key, v in forc.items() if key.endswith(" 1") for i in range(replicate)}
            qout = q.run(q.gather(series, dep), chunk=chunk)
            errs[qsec] = q.segment_errors(replicate)
            kname = kname or q.kernel_name()
            for j, (i, k) in enumerate(q.rows):
                for r, full in enumerate(q.save):
                    key = "%s%d_%s" % (qsec, k, full.split("_", 1)[1])
                    if key not in ts:
                        ts[key] = np.empty((qout.shape[0], replicate))
                    ts[key][:, i] = qout[:, r, j]
    return ts, errs, kname


def run_dropin(case, time_unit="us", device=0):
    """The reference's dispatch order through the drop-in activities, one section per call."""
    si = CASES.siminfo_for(case)
    si["time_unit"] = time_unit
    ts = dict(CASES.forcings(case))
    activities.DEVICE = device
    acts = dict(activities.ACTIVITIES[case["operation"]])
    errs = CASES.run_segment(case["operation"], copy.deepcopy(case["tables"]), ts, si, acts)
    return ts, errs


def rel_err(got, ref):
    """max elementwise |got-ref| / max(|ref|, 1e-4*max|ref|, tiny), with NaN==NaN and +-1e30 sentinels exact."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if got.shape != ref.shape:
        return np.inf
    nan_r, nan_g = np.isnan(ref), np.isnan(got)
    if not np.array_equal(nan_r, nan_g):
        return np.inf
    sent = np.abs(ref) > 1e29
    if not np.array_equal(got[sent], ref[sent]):
        return np.inf
    ok = ~(nan_r | sent)
    if not ok.any():
        return 0.0
    scale = np.abs(ref[ok]).max()
    den = np.maximum(np.abs(ref[ok]), max(1e-4 * scale, 1e-300))
    return float((np.abs(got[ok] - ref[ok]) / den).max())


def compare(ts_ref, ts_got, exact, seg=None):
    """-> list of (name, error) violations over the series present in ts_got that the oracle also produced."""
    bad = []
    for k, g in ts_got.items():
        if k not in ts_ref:
            bad.append((k, "not produced by the oracle"))
            continue
        r = ts_ref[k]
        g = np.asarray(g)
        if g.ndim == 2:
            cols = range(g.shape[1]) if seg is None else [seg]
        else:
            cols = [None]
        for c in cols:
            gc = g if c is None else g[:, c]
            n = min(len(r), len(gc))
            if exact:
                if not np.array_equal(gc[:n], r[:n], equal_nan=True):
                    bad.append((k, "bits differ (max rel %.3g)" % rel_err(gc[:n], r[:n])))
                    break
            else:
                e = rel_err(gc[:n], r[:n])
                if not e <= RTOL:
                    bad.append((k, e))
                    break
    return bad


def f32_ulp_diff(got32, ref64):
    """Largest distance, in float32 units in the last place, between float32 results and the float32 cast of the fp64
    reference (what save_timeseries stores, utilities.py:313); NaN positions must agree."""
    got32 = np.asarray(got32)
    assert got32.dtype == np.float32
    with np.errstate(over="ignore"):
        ref32 = np.asarray(ref64, dtype=np.float64).astype(np.float32)
    if not np.array_equal(np.isnan(got32), np.isnan(ref32)):
        return np.inf
    ok = ~np.isnan(ref32)
    a = got32[ok].view(np.int32).astype(np.int64)
    b = ref32[ok].view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)  # sign-magnitude -> monotone integer line
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return int(np.abs(a - b).max()) if a.size else 0


def check_metric_batch_kernels(names=("p_metric_pack", "i001_metric", "cbp_fullchain_metric", "i_chain_metric"), replicate=3):
    """siminfo['units'] == 2 folded into the kernel built for the batch (StaticFlags<.., OPTS>: the conversions of the stored
    series are compiled in, the English-unit batches carry none of them) gives the bits of the general kernel, which reads the
    option from the launch record."""
    for name in names:
        case = CASES.build_cases()[name]
        a, ea, ka = run_fused(case, replicate=replicate, chunk=200, jit=False)
        b, eb, kb = run_fused(case, replicate=replicate, chunk=200, jit=True)
        assert ",jit>" in kb and "batch-opts=0x2" in kb and ",jit>" not in ka, (ka, kb)
        assert a.keys() == b.keys()
        for k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), (name, k)
        for k in ea:
            assert np.array_equal(ea[k], eb[k]), (name, k)


def check_shared_met_mode(n=75, jit=None, nst=5, table=False):
    """hsp2g_set_shared_forcing: a few met stations + per-segment station index and MFACTOR give bit-identical results to
    handing every segment its own series * MFACTOR arrays (what get_timeseries builds, utilities.py:274-290); generic and
    compile-time-specialised kernels, with and without divergence ordering."""
    cal = Calendar("1976-02-01", "1976-03-10", 60, "us")
    steps = cal.steps
    # per-lane gathers from the station table; HSP2G_MET_TABLE=1 and nst <= 32: one bulk copy of the table per warp-interval
    old = os.environ.get("HSP2G_MET_TABLE")
    os.environ["HSP2G_MET_TABLE"] = "1" if table else "0"
    try:
        _check_shared_met_mode(n, jit, nst, table, cal, steps)
    finally:
        if old is None:
            del os.environ["HSP2G_MET_TABLE"]
        else:
            os.environ["HSP2G_MET_TABLE"] = old


def _check_shared_met_mode(n, jit, nst, table, cal, steps):
    rng = np.random.default_rng(3)
    ens = synth.Ensemble("PERLND", n, seed=21, sections=("SNOW", "PWATER"))
    names = list(synth.PERLND_FORCINGS)
    s0 = synth.shared_series("PERLND", cal, 0, steps)
    series = np.empty((steps, len(names), nst))
    for k in range(nst):  # stations differ by a temperature offset and a rain factor
        for r, nm in enumerate(names):
            a = s0[nm].copy()
            if nm == "AIRTMP":
                a = a + (k % 5 - 2) * 1.5
            elif nm == "DTMPG":
                a = (s0["AIRTMP"] + (k % 5 - 2) * 1.5) - 5.0
            elif nm == "PREC":
                a = a * (1.0 + 0.05 * (k % 7))
            series[:, r, k] = a
    station = rng.integers(0, nst, n).astype(np.int32)
    mfac = np.ones((len(names), n))
    mfac[names.index("PREC")] = 1.0 + 0.2 * ens.u_prec
    mfac[names.index("PETINP")] = 1.0 + 0.1 * ens.u_pet
    expanded = series[:, :, station] * mfac[None, :, :]
    store = ens.store(cal)
    for save in ("all", ["SNOW_" + k for k in test10.SAVE_DEFAULT["SNOW"]] +
                 ["PWATER_" + k for k in test10.SAVE_DEFAULT["PWATER"]]):
        with LandBatch(store, cal, names, save, jit=False) as b:
            ref = b.run(expanded, chunk=300)
            ref_state = b.get_state()
        for order in (None, "flags", np.argsort(station, kind="stable")):
            with LandBatch(store, cal, names, save, order=order, jit=jit, layout="tiled" if order is None else "plain") as b:
                b.set_shared_forcing(series, station, mfac)
                got = b.run_shared(chunk=300)
                assert "met=shared" in b.kernel_name()
                assert not jit or ("met=shared/table" in b.kernel_name()) == (table and nst <= 32), b.kernel_name()
                assert jit is None or (",jit>" in b.kernel_name()) == bool(jit), b.kernel_name()
                assert np.array_equal(b.get_state(), ref_state, equal_nan=True)
            assert np.array_equal(got, ref, equal_nan=True)


def synthetic_sources():
    rng = np.random.default_rng(11)
    t0 = np.datetime64("1976-01-01", "s")

    def mk(dsn, step_s, tcode, tstep, n, scale, offset=0.0):
        return WdmSeries(dsn, {"TCODE": tcode, "TSSTEP": tstep}, t0 + np.arange(n) * np.timedelta64(step_s, "s"),
                         (rng.random(n) * scale + offset).astype(np.float32).astype(np.float64))

    days = 120
    src = {"P1": mk(1, 3600, 3, 1, 24 * days, 0.05), "P2": mk(2, 3600, 3, 1, 24 * days, 0.08),
           "T1": mk(3, 7200, 3, 2, 12 * days, 30.0, 10.0), "T2": mk(4, 7200, 3, 2, 12 * days, 25.0, 5.0),
           "E1": mk(5, 86400, 4, 1, days, 0.2), "W1": mk(6, 86400, 4, 1, days, 100.0),
           "S1": mk(7, 7200, 3, 2, 12 * days, 40.0), "D1": mk(8, 86400, 4, 1, days, 20.0, 5.0)}
    src["P1"].values[rng.random(24 * days) < 0.8] = 0.0
    src["P2"].values[rng.random(24 * days) < 0.8] = 0.0
    return src


def _rows(prec, pf, temp, pet_f):
    return [SourceRow(prec, pf, "SAME", "PREC"), SourceRow(temp, 1.0, "SAME", "AIRTMP"), SourceRow("E1", pet_f, "DIV", "PETINP"),
            SourceRow("W1", 1.0, "DIV", "WINMOV"), SourceRow("S1", 1.0, "DIV", "SOLRAD"), SourceRow("D1", 1.0, "SAME", "DTMPG")]


def check_metstore_shared():
    """40 segments on 2 rain gages x 2 temperature gages with per-segment PREC MFACTORs: 4 + 2 stations, factors on the
    device where that is exact, and the shared-met kernels reproduce the per-segment-array kernels bit for bit."""
    cal = Calendar("1976-01-10", "1976-03-20", 60, "us")
    src = synthetic_sources()
    n = 40
    rng = np.random.default_rng(5)
    rows = []
    for i in range(n):
        rows.append(_rows("P1" if i % 2 else "P2", float(1.0 + 0.01 * (i % 7)), "T1" if (i // 2) % 2 else "T2",
                          0.7 if i % 5 else 0.75))
    rows[3].append(SourceRow("P2", 0.5, "SAME", "PREC"))  # a second row feeding the same target is summed on the host
    names = list(synth.PERLND_FORCINGS)
    calls = []
    ms = MetStore.build(cal, names, rows, lambda k: (calls.append(k), src[k])[1])
    assert sorted(calls) == sorted(src)  # every source read once
    assert ms.n_stations == 2 * 2 * 2 + 1 and ms.series.shape == (cal.steps, 6, ms.n_stations)
    j = names.index("PREC")
    assert ms.mfactor[j, 5] == 1.0 + 0.01 * 5 and ms.mfactor[j, 3] == 1.0  # the summed target is baked into its own station
    assert np.all(ms.mfactor[names.index("PETINP")] == 1.0)  # DIV: (x * MFACTOR) * ratio is formed on the host
    full = ms.expand()
    # the reference's order of operations, spelled out for two segments
    for i in (3, 5):
        acc = {}
        for r in rows[i]:
            s = src[r.svolno]
            v = s.values * r.mfactor if r.mfactor != 1.0 else s.values
            t = resample(s.times, v, s.step, r.tmemn, r.tran, cal, s.tcode)
            acc[r.tmemn] = acc[r.tmemn] + t if r.tmemn in acc else t
        for jj, nm in enumerate(names):
            assert np.array_equal(full[:, jj, i], acc[nm]), (i, nm)
    ens = synth.Ensemble("PERLND", n, seed=3, sections=("SNOW", "PWATER"))
    store = ens.store(cal)
    save = ["SNOW_" + k for k in test10.SAVE_DEFAULT["SNOW"]] + ["PWATER_" + k for k in test10.SAVE_DEFAULT["PWATER"]]
    with LandBatch(store, cal, names, save) as b:
        ref = b.run(full, chunk=500)
        ref_state = b.get_state()
    with LandBatch(store, cal, names, save, order="flags") as b:
        ms.apply(b)
        got = b.run_shared(chunk=500)
        assert "met=shared" in b.kernel_name()
        assert np.array_equal(b.get_state(), ref_state, equal_nan=True)
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.nanmax(ref[:, save.index("SNOW_PACKF")]) > 0 and np.nanmax(ref[:, save.index("PWATER_SURO")]) > 0
    with LandBatch(store, cal, names[:5] + ["CLOUD"], save) as b:
        try:
            ms.apply(b)
        except ValueError:
            pass
        else:
            raise AssertionError("a store built for other forcings was accepted")


class FrameSink:
    """Records write_ts calls: the SupportsWriteTS surface as IOManager exposes it (hsp2io/io.py:58-66)."""

    def __init__(self):
        self.calls, self.bytes = {}, 0

    def write_ts(self, data_frame, save_columns, category, operation=None, segment=None, activity=None, compress=True,
                 outstep=2):
        self.calls[(operation, segment, activity)] = (data_frame, list(save_columns))
        self.bytes += data_frame.to_numpy().nbytes


def write_gage_wdm(path, start, days, gages):
    """A synthetic WDM file of `gages` met stations built from test10's 1976 met (datasets 100g+1 .. 100g+6: hourly PREC,
    2-hourly ATMP, daily EVAP, daily WIND, 2-hourly SOLR, daily DEWP), written with tests/wdm_writer.py."""
    from tests.wdm_writer import write_wdm

    start = np.datetime64(start, "s")
    cal = Calendar(start, start + np.timedelta64(days, "D"), 60, "ns")
    m = synth.shared_series("PERLND", cal, 0, cal.steps)
    sets = []
    for g in range(gages):
        hourly = lambda a: a
        two = lambda a: a.reshape(-1, 2)
        day = lambda a: a.reshape(-1, 24)
        series = {1: (3, 1, hourly(m["PREC"]) * (1.0 + 0.04 * g)),
                  2: (3, 2, two(m["AIRTMP"])[:, 0] + 0.8 * (g - gages / 2)),
                  3: (4, 1, day(m["PETINP"]).sum(axis=1) / 0.7),
                  4: (4, 1, day(m["WINMOV"]).sum(axis=1)),
                  5: (3, 2, two(m["SOLRAD"]).sum(axis=1)),
                  6: (4, 1, day(m["DTMPG"])[:, 0] + 0.8 * (g - gages / 2))}
        for k, (tcode, tstep, v) in series.items():
            v = np.asarray(v, dtype=np.float32)
            # yearly groups; long hourly records are cut into blocks, and runs of zeros are stored compressed
            blocks, i = [], 0
            while i < len(v):
                j = i
                if v[i] == 0.0:
                    while j < len(v) and v[j] == 0.0 and j - i < 30000:
                        j += 1
                    blocks.append((v[i:j], True))
                else:
                    while j < len(v) and v[j] != 0.0 and j - i < 400:
                        j += 1
                    blocks.append((v[i:j], False))
                i = j
            year_end = (start.astype("datetime64[Y]") + 1).astype("datetime64[s]")  # a group is always full: pad with TFILL
            per = np.timedelta64(3600 * tstep if tcode == 3 else 86400 * tstep, "s")
            missing = int((year_end - start) / per) - len(v)
            if missing > 0:
                blocks.append((np.full(missing, -999.0, dtype=np.float32), True))
            sets.append({"dsn": 100 * g + k, "tcode": tcode, "tstep": tstep, "tgroup": 6, "groups": [(start, blocks)]})
    write_wdm(path, sets)


def check_watershed(tmpdir, n=48, gages=4, days=40, sample=(0, 7, 21), chunk=480):
    """BASELINE.json configs[4] in miniature: a watershed of `n` PERLND segments running the full chain with per-segment option
    flags draws its forcings from a WDM file of `gages` met stations (WdmFile -> MetStore -> shared-met kernels, no per-segment
    forcing arrays), and its SAVEd series leave as the float32 frames save_timeseries builds (results.save_batch).  Sampled
    segments must equal the oracle run on get_timeseries-style arrays, after the float32 cast, bit for bit."""
    import pandas as pd

    from hspsquared_b200.results import save_batch
    from hspsquared_b200.wdm import WdmFile

    path = os.path.join(str(tmpdir), "gages.wdm")
    start = "1976-01-01"
    write_gage_wdm(path, start, days + 2, gages)
    stop = np.datetime64(start, "s") + np.timedelta64(days, "D")
    cal = Calendar(start, stop, 60, "us")
    wdm = WdmFile(path)
    assert len(wdm.dsns) == 6 * gages
    ens = synth.Ensemble("PERLND", n, seed=99, base="cbp", options=("RTOPFG", "UZFG", "SNOPFG", "ICEFG", "TSOPFG"))
    names = list(synth.FULLCHAIN_FORCINGS)
    member = {"PREC": (1, "SAME"), "GATMP": (2, "SAME"), "PETINP": (3, "DIV"), "WINMOV": (4, "DIV"), "SOLRAD": (5, "DIV"),
              "DTMPG": (6, "SAME")}
    rows = []
    for i in range(n):
        g = (i * 7) % gages
        rows.append([SourceRow("TS%03d" % (100 * g + member[nm][0]),
                               (1.0 + 0.02 * (i % 5)) if nm == "PREC" else 0.7 if nm == "PETINP" else 1.0, member[nm][1], nm)
                     for nm in names])
    ms = MetStore.build(cal, names, rows, lambda k: wdm.series(int(k[2:])))
    assert ms.n_stations == gages
    store = ens.store(cal)
    with LandBatch(store, cal, names, "all", order="flags") as b:
        ms.apply(b)
        b.set_output_dtype(np.float32)
        out = b.run_shared(chunk=chunk)
        saved = b.save
        assert "met=shared" in b.kernel_name()
    acts = ("ATEMP", "SNOW", "PWATER", "SEDMNT", "PSTEMP", "PWTGAS")
    save_tables = {}
    for full in saved:
        a, k = full.split("_", 1)
        save_tables.setdefault(a, {})[k] = 1
    assert set(save_tables) == set(acts) and len(saved) == 84
    tindex = pd.date_range(start, periods=cal.steps + 1, freq="60min")[1:]  # main.py:93-95
    sink = FrameSink()
    seg_names = ["P%04d" % (i + 1) for i in range(n)]
    save_batch(sink, {"tindex": tindex}, "PERLND", seg_names, save_tables, saved, out)
    assert len(sink.calls) == n * len(acts)
    for i in sample:
        f = ms.expand(segments=[i])[:, :, 0]
        ts = {nm: np.ascontiguousarray(f[:, j]) for j, nm in enumerate(names)}
        si = {"start": cal.start, "stop": cal.stop, "delt": 60, "units": 1, "steps": cal.steps, "time_unit": "us"}
        CASES.run_segment("PERLND", ens.tables_of(i), ts, si, ORACLE_ACTS)
        for a in acts:
            df, cols = sink.calls[("PERLND", seg_names[i], a)]
            assert list(df.columns) == sorted(save_tables[a]) and sorted(cols) == sorted(save_tables[a])
            for k in df.columns:
                with np.errstate(over="ignore"):
                    want = ts[k].astype(np.float32)
                got = df[k].to_numpy()
                assert got.dtype == np.float32
                if GPU_EXACT:
                    assert np.array_equal(got, want, equal_nan=True), (i, a, k)
                else:
                    assert f32_ulp_diff(got, ts[k]) <= 2, (i, a, k)
    return sink


def check_qual_chain_on_device(n=70, chunk=300, order=None, chain="gather", layout="tiled", jit=None):
    """Land batch -> PQUAL rows with every buffer device-resident gives the same bits as feeding the constituents from the land
    batch's host-side outputs (QualBatch.gather), with and without a divergence-aware device order of the land segments.
    chain="gather": hsp2g_gather_* fills a scratch buffer (tile-major); chain="linked": the rows read the land batch's buffers in
    place (hsp2g_run_device_linked), in either series layout."""
    import ctypes as C

    from hspsquared_b200.engine import QualBatch, tile_series, untile_series
    from hspsquared_b200.qualstore import PQUAL_INPUTS

    cal = Calendar("1976-02-10", "1976-03-25", 60, "us")
    steps = cal.steps
    ens = synth.Ensemble("PERLND", n, seed=31, base="cbp", options=("RTOPFG", "UZFG"))
    names = list(synth.FULLCHAIN_FORCINGS)
    f = synth.forcing_chunk(ens, cal, 0, steps, names=names)
    need = ["PWATER_SURO", "PWATER_IFWO", "PWATER_AGWO", "PWATER_PERO", "SEDMNT_WSSD", "SEDMNT_SCRSD", "SNOW_PACKF"]
    qn = list(PQUAL_INPUTS[:7])
    ucis = [test10.pqual_tables() for _ in range(n)]
    lib = _lib.load()
    # reference path: host round trip
    with LandBatch(ens.store(cal), cal, names, need) as land:
        lo = land.run(f, chunk=chunk)
    series = {full.split("_", 1)[1]: lo[:, r, :] for r, full in enumerate(need)}
    series["PREC"] = f[:, names.index("PREC"), :]
    with QualBatch("PERLND", ucis, cal, qn, "all") as q:
        want = q.run(q.gather(series), chunk=chunk)
        want_rows = list(q.rows)
    # device-resident chain
    if chain == "gather":
        layout = "tiled"
    with LandBatch(ens.store(cal), cal, names, need, order=order, layout=layout, jit=jit) as land, \
            QualBatch("PERLND", ucis, cal, qn, "all", land=land, chain=chain, layout=layout, jit=jit) as q:
        assert q.chain == chain
        got = np.empty((steps, q.ns, q.n))
        dev = []

        def dalloc(nbytes):
            p = lib.hsp2g_device_alloc(0, nbytes)
            assert p
            dev.append(p)
            return p

        for t0 in range(0, steps, chunk):
            m = min(chunk, steps - t0)
            fc = f[t0:t0 + m][:, land._fperm, :]
            if land.order is not None:
                fc = fc[:, :, land.order]
            if layout == "tiled":
                ft = tile_series(fc)
            else:
                ft = np.zeros((m, land.nf, land.npad))
                ft[:, :, :land.n] = fc
            d_f, d_lo = dalloc(ft.nbytes), dalloc(land.npad * m * land.ns * 8)
            d_qo = dalloc(q.npad * m * q.ns * 8)
            _lib.check(lib.hsp2g_memcpy_h2d(0, d_f, ft.ctypes.data, ft.nbytes))
            land.run_device(t0, m, d_f, d_lo)
            if chain == "gather":
                q.run_device_from(t0, m, d_f, d_lo, dalloc(q.npad * m * q.nf * 8), d_qo)
            else:
                q.run_device_linked_from(t0, m, d_f, d_lo, d_qo)
                assert "in=linked-rows" in q.kernel_name() or not q.jit, q.kernel_name()
            land.sync()
            q.sync()
            if layout == "tiled":
                ho = np.empty((q.ntiles, m, q.ns, 32))
                _lib.check(lib.hsp2g_memcpy_d2h(0, ho.ctypes.data, d_qo, ho.nbytes))
                u = untile_series(ho, q.n)
            else:
                ho = np.empty((m, q.ns, q.npad))
                _lib.check(lib.hsp2g_memcpy_d2h(0, ho.ctypes.data, d_qo, ho.nbytes))
                u = ho[:, :, :q.n]
            u = u[:, q._sinv, :]  # device row order -> the caller's
            got[t0:t0 + m] = u if q._inv is None else u[:, :, q._inv]
        for p in dev:
            lib.hsp2g_device_free(0, C.c_void_p(p))
        rows = list(q.rows)
    if chain == "linked":  # one row per (land device segment incl. the padding lanes, constituent): pick the real ones
        assert len(rows) == land.npad * 3 and sum(1 for sgm, _k in rows if sgm < 0) == (land.npad - n) * 3
        at = {r: j for j, r in enumerate(rows) if r[0] >= 0}
        got = got[:, :, [at[r] for r in want_rows]]
    assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True)
    assert np.nanmax(want[:, q.save.index("PQUAL_SOQUAL"), :]) > 0


def check_abi_errors_and_edges():
    """Error behaviour of the C ABI (status + hsp2g_last_error, never a crash or a silent default) and the smallest shapes:
    one segment, one interval, one interval per chunk, an empty SAVE set."""
    import ctypes as C

    lib = _lib.load()
    h = _lib._h()

    def err():
        return lib.hsp2g_last_error().decode()

    assert lib.hsp2g_batch_create(0, 7, 4, 6, 60.0, C.byref(h)) != 0 and "operation" in err()
    assert lib.hsp2g_batch_create(0, 0, 0, 6, 60.0, C.byref(h)) != 0 and "n_seg" in err()
    assert lib.hsp2g_batch_create(0, 0, 4, 6, 0.0, C.byref(h)) != 0 and "delt" in err()
    assert lib.hsp2g_batch_create(0, 0, 4, 64, 60.0, C.byref(h)) != 0 and "activity mask" in err()  # IWATER on PERLND
    assert lib.hsp2g_batch_create(0, 0, 4, 512 | 4, 60.0, C.byref(h)) != 0 and "PQUAL" in err()  # constituents + hydrology
    assert lib.hsp2g_batch_create(0, 0, 4, 2, 720.0, C.byref(h)) != 0 and "ICEFLG" in err()  # SNOW.py:85
    assert lib.hsp2g_batch_create(0, 0, 4, 6, 60.0, None) != 0
    assert lib.hsp2g_batch_create(0, 0, 3, 6, 60.0, C.byref(h)) == 0 and h.value
    out = np.zeros(3 * 24 * 2)
    assert lib.hsp2g_run_host(h, 0, 1, out.ctypes.data, out.ctypes.data) != 0 and "hsp2g_set_params" in err()
    assert lib.hsp2g_set_params(h, None, 3) != 0
    assert lib.hsp2g_set_monthly(h, 999, out.ctypes.data_as(C.POINTER(C.c_double)), 3) != 0 and "out of range" in err()
    assert lib.hsp2g_set_schedule(h, -1) != 0
    assert lib.hsp2g_set_output_dtype(h, 5) != 0
    bad = np.array([9999], dtype=np.int32)
    assert lib.hsp2g_set_save(h, 1, bad.ctypes.data_as(C.POINTER(C.c_int32))) != 0
    assert lib.hsp2g_batch_destroy(h) == 0
    assert lib.hsp2g_run_host(None, 0, 1, None, None) != 0 and "null" in err()
    # ---- smallest shapes against the oracle
    case = dict(CASES.build_cases()["p001_snow_pwater"])
    for steps, chunk in ((1, None), (2, 1), (25, 1)):
        c = dict(case, steps=steps)
        ref, _ = run_oracle(c)
        got, _, _ = run_fused(c, chunk=chunk)
        assert not compare(ref, got, exact=GPU_EXACT), (steps, chunk)
    si = CASES.siminfo_for(dict(case, steps=30))
    cal = Calendar(si["start"], si["stop"], 60, "us")
    forc = {k: v[:30] for k, v in CASES.forcings(case).items()}
    names = [n for n in forc if n in _lib.index("forcings")]
    store = SegmentStore.from_tables("PERLND", [copy.deepcopy(case["tables"])], cal)
    with LandBatch(store, cal, names, []) as b:  # nothing saved: the state still advances
        o = b.run(pack_forcings(forc, names, 1), chunk=7)
        assert o.shape == (30, 0, 1)
        st_none = b.get_state()
    with LandBatch(store, cal, names, "all") as b:
        b.run(pack_forcings(forc, names, 1))
        assert np.array_equal(b.get_state(), st_none, equal_nan=True)
    with LandBatch(store, cal, names, "all", layout="tiled") as b:
        try:
            b.run_host(25, np.zeros((1, 10, len(names), 32)))  # 25 + 10 > 30 intervals of calendar
        except _lib.Hsp2gError as e:
            assert "outside the calendar" in str(e)
        else:
            raise AssertionError("a chunk past the calendar was accepted")


def check_option_word_kernels():
    """Kernels built for one batch (hspsquared_b200/jit.py: activity mask, row sets, element types, layout and the batch-wide
    option word fixed at compile time) against (i) the same batch with one segment's word changed, whose kernel tests the
    options at run time, and (ii) the library's own general kernels: every segment must come out bit for bit the same, outputs
    and state.  The words are test10's PERLND 1 / IMPLND 1 and the CBP full-chain segment's -- any word gets its kernel."""
    cal = Calendar("1976-01-20", "1976-03-05", 60, "us")
    pd_ = ["SNOW_" + k for k in test10.SAVE_DEFAULT["SNOW"]] + ["PWATER_" + k for k in test10.SAVE_DEFAULT["PWATER"]]
    id_ = ["SNOW_" + k for k in test10.SAVE_DEFAULT["SNOW"]] + ["IWATER_" + k for k in test10.SAVE_DEFAULT["IWATER"]]
    P6, FC = list(synth.PERLND_FORCINGS), list(synth.FULLCHAIN_FORCINGS)
    jobs = [("PERLND", dict(sections=("SNOW", "PWATER")), P6, pd_, "IFRDFG", "plain"),
            ("PERLND", dict(sections=("SNOW", "PWATER")), P6, "all", "IFRDFG", "tiled"),
            ("PERLND", dict(sections=("SNOW", "PWATER")), P6, ["PWATER_SURO", "PWATER_IFWO", "PWATER_AGWO"], "IFRDFG", "plain"),
            ("PERLND", dict(base="cbp"), FC, "all", "RTOP1", "plain"),
            ("PERLND", dict(base="cbp"), FC, list(cbp10()), "UZFG", "tiled"),
            ("IMPLND", dict(sections=("SNOW", "IWATER")), P6, id_, "RTOP1", "plain"),
            ("IMPLND", dict(sections=("SNOW", "IWATER", "SOLIDS")), P6, id_, "RTOP1", "tiled")]
    for op, kw, names, save, bit, layout in jobs:
        ens = synth.Ensemble(op, 70, seed=9, **kw)
        f = synth.forcing_chunk(ens, cal, 0, cal.steps, names=names)
        word = int(ens.store(cal).flags[0])
        with LandBatch(ens.store(cal), cal, names, save, layout=layout, jit=True) as b:
            a = b.run(f, chunk=400)
            assert b.jit_error is None, b.jit_error
            assert "options=%#x" % word in b.kernel_name() and ",jit>" in b.kernel_name() and layout in b.kernel_name(), b.kernel_name()
            sa = b.get_state()
        st = ens.store(cal)
        st.flags[-1] ^= np.uint64(1) << np.uint64(st.FB[bit])
        with LandBatch(st, cal, names, save, layout=layout, jit=True, order=None) as b:
            g = b.run(f, chunk=400)
            # one segment differs in one option bit: that bit is read per segment, the other 61 stay fixed at compile time
            assert "(1 bits per segment)" in b.kernel_name() and ",jit>" in b.kernel_name(), b.kernel_name()
            sg = b.get_state()
        with LandBatch(ens.store(cal), cal, names, save, layout=layout, jit=False) as b:
            c = b.run(f, chunk=400)
            assert "jit" not in b.kernel_name(), b.kernel_name()
            sc = b.get_state()
        tag = (op, kw, layout)
        assert np.array_equal(a[:, :, :-1], g[:, :, :-1], equal_nan=True), tag
        assert np.array_equal(sa[:, :-1], sg[:, :-1], equal_nan=True), tag
        assert np.array_equal(a, c, equal_nan=True), tag
        assert np.array_equal(sa, sc, equal_nan=True), tag


def cbp10():
    """the ten series a CBP land-segment UCI hands its river model (tests/land_spec/hwmA51800.uci:493-502)"""
    return ("PWATER_SURO", "PWATER_IFWO", "PWATER_AGWO", "SEDMNT_SOSED", "PWTGAS_SOHT", "PWTGAS_IOHT", "PWTGAS_AOHT",
            "PWTGAS_SODOXM", "PWTGAS_IODOXM", "PWTGAS_AODOXM")


def check_period_aggregation(n=40, days=75):
    """hsp2g_aggregate_f32 against pandas on the float32 frames (what IOManager.write_ts does for outstep 3 / 4 / 5,
    hsp2io/io.py:74-94): last / sum / mean per day, month and year, NaN series included, bit for bit."""
    import ctypes as C

    import pandas as pd

    from hspsquared_b200.results import aggregate_host_series, period_bins, save_batch_aggregated

    start = np.datetime64("1976-01-20T00:00:00")
    cal = Calendar(start, start + np.timedelta64(days, "D"), 60, "us")
    ens = synth.Ensemble("PERLND", n, seed=4, sections=("SNOW", "PWATER"))
    names = list(synth.PERLND_FORCINGS)
    with LandBatch(ens.store(cal), cal, names, "all") as b:
        b.set_output_dtype(np.float32)
        out = b.run(synth.forcing_chunk(ens, cal, 0, cal.steps, names=names), chunk=600)
        saved = b.save
    tindex = pd.date_range(pd.Timestamp(str(start)), periods=cal.steps + 1, freq="60min")[1:]
    assert np.isnan(out[:, saved.index("SNOW_RDENPF")]).any()  # a series with NaN stretches (no pack)
    for outstep in (3, 4, 5):
        bins, labels = period_bins(tindex, outstep)
        agg = aggregate_host_series(out, bins, len(labels))
        for i in (0, n // 2, n - 1):
            df = pd.DataFrame(out[:, :, i], index=tindex, columns=[s.split("_", 1)[1] + "." + s.split("_", 1)[0] for s in saved])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")  # pandas 3: origin= has no effect for calendar frequencies (it says so)
                # daily: the 24-hour Tick anchored at the first label, which is what 'D' with origin='start' means in the pandas
                # the reference pins (results.period_bins)
                r = df.resample({3: "24h", 4: "ME", 5: "YE"}[outstep], origin="start")
            for stat, want in enumerate((r.last(), r.sum(), r.mean())):
                assert want.index.equals(labels)
                w = want.to_numpy()
                assert w.dtype == np.float32
                assert np.array_equal(agg[:, :, stat, i], w, equal_nan=True), (outstep, stat, i)
    # the same statistics delivered by hsp2g_run_host itself (hsp2g_set_output_periods), chunks cut at period ends
    bins, labels = period_bins(tindex, 3)
    want = aggregate_host_series(out, bins, len(labels))
    with LandBatch(ens.store(cal), cal, names, "all", order="flags") as b:
        got = b.run_periods(synth.forcing_chunk(ens, cal, 0, cal.steps, names=names), bins, periods_per_chunk=9)
        try:
            b.set_output_dtype(np.float32)
            _lib.check(b.lib.hsp2g_set_output_periods(b.h, len(bins), bins.ctypes.data_as(C.POINTER(C.c_int32))))
            b.run_chunk(0, np.zeros((30, b.nf, b.n)), np.empty((30, b.ns, b.n), dtype=np.float32))  # 30 intervals: not whole days
        except _lib.Hsp2gError as e:
            assert "period boundaries" in str(e)
        else:
            raise AssertionError("a chunk that splits a period was accepted")
    assert np.array_equal(got, want, equal_nan=True)
    # frame layout of the merged result (io.py:79-81)
    sink = FrameSinkBackend()
    bins, labels = period_bins(tindex, 3)
    agg = aggregate_host_series(out, bins, len(labels))
    tabs = {"SNOW": {"PACKF": 1, "RDENPF": 1, "DULL": 0}, "PWATER": {"SURO": 1}}
    save_batch_aggregated(sink, labels, "PERLND", ["P%03d" % k for k in range(n)], tabs, saved, agg)
    df = sink.calls[("PERLND", "P003", "SNOW")]
    assert list(df.columns) == ["PACKF_last", "RDENPF_last", "PACKF_sum", "RDENPF_sum", "PACKF_aver", "RDENPF_aver"]
    assert len(sink.calls) == 2 * n and df.index.equals(labels)


class FrameSinkBackend:
    def __init__(self):
        self.calls = {}

    def write_ts(self, data_frame, category, operation=None, segment=None, activity=None, *a, **k):
        self.calls[(operation, segment, activity)] = data_frame


def check_doe(doe_list, device=0):
    """doe.run_doe (mainDoE.py:19-140 as ONE batch of (run, segment) rows) against the oracle run by run: every saved series
    and every error vector of every (run, segment) equals the oracle on that run's tables."""
    from hspsquared_b200 import doe as DOE

    case = CASES.build_cases()["p001_snow_pwater"]
    case = dict(case, steps=24 * 100)
    si = CASES.siminfo_for(case)
    si["time_unit"] = "us"
    forc = CASES.forcings(case)
    t2 = copy.deepcopy(case["tables"])
    t2["PWATER"]["PARAMETERS"]["LZSN"] = 6.0
    segments = {"P001": case["tables"], "P002": t2}
    forcings = {"P001": forc, "P002": {k: v * (1.1 if k == "PREC" else 1.0) for k, v in forc.items()}}
    results, errors = DOE.run_doe("PERLND", segments, forcings, si, doe_list, chunk=700, device=device)
    assert set(results) == {"1", "2", "5", "10", "12", "13", "14", "15"}
    runlist = DOE.make_runlist(doe_list)
    for run in results:
        for seg in segments:
            tabs = DOE.tables_for_run("PERLND", seg, segments[seg], runlist[run])
            ts = dict(forcings[seg])
            e = CASES.run_segment("PERLND", tabs, ts, dict(si), ORACLE_ACTS)
            for k, g in results[run][seg].items():
                if GPU_EXACT:
                    assert np.array_equal(g, ts[k], equal_nan=True), (run, seg, k)
                else:
                    assert rel_err(g, ts[k]) <= RTOL, (run, seg, k)
            for a in e:
                assert np.array_equal(errors[run][seg][a], e[a]), (run, seg, a)
    # the runs really differ
    assert not np.array_equal(results["1"]["P001"]["SURO"], results["2"]["P001"]["SURO"])
    # ICEFG 0 / 1 (runs 14 / 15): the ice content of the pack grows only with ICEFG on (SNOW.py:461-485), and PWATER's
    # frozen-ground infiltration factor follows it (PWATER.py:295-296)
    assert results["15"]["P001"]["PACKI"].max() > results["14"]["P001"]["PACKI"].max()
    assert not np.array_equal(results["14"]["P001"]["INFFAC"], results["15"]["P001"]["INFFAC"])
    assert np.array_equal(results["12"]["P002"]["CEPS"], results["14"]["P002"]["CEPS"])  # untouched segment: same run
    return results


# ---- a dispatcher shaped like hsp2.main (main.py:91-275) for boxes without the reference tree: the same local variable
# names, the same call, so that the drop-in activities find it the way they find main() (lazybatch.main_context)
_DISPATCH_SERIES = {}


def get_timeseries(io_manager, ext_sourcesdd, siminfo):  # noqa: D401 -- stands in for utilities.get_timeseries
    return {k: v.copy() for k, v in _DISPATCH_SERIES[ext_sourcesdd].items()}


def _dispatch_like_main(io_manager, uci_obj, registry):
    import pandas as pd

    opseq, uci, ddext_sources, siminfo = uci_obj["opseq"], uci_obj["uci"], uci_obj["ddext_sources"], uci_obj["siminfo"]
    ddlinks = {}
    results, errors_of = {}, {}
    for _, operation, segment, delt in opseq.itertuples():
        siminfo["delt"] = delt
        ts = get_timeseries(io_manager, ddext_sources[(operation, segment)], siminfo)
        flags = uci[(operation, "GENERAL", segment)]["ACTIVITY"]
        for activity, function in registry[operation].items():
            if (activity in flags) and (not flags[activity]):
                continue
            if (operation, activity, segment) not in uci:
                continue
            ui = uci[(operation, activity, segment)]
            if operation == "PERLND" and activity in ("SEDMNT", "PWTGAS"):
                ui["PARAMETERS"]["CSNOFG"] = uci[(operation, "PWATER", segment)]["PARAMETERS"]["CSNOFG"]
            if operation == "PERLND" and activity == "PSTEMP":
                ui["PARAMETERS"]["AIRTFG"] = flags["ATEMP"]
            errors, errmessages = function(io_manager, siminfo, ui, ts)
            errors_of[(operation, segment, activity)] = np.asarray(errors)
        results[(operation, segment)] = ts
    del pd
    return results, errors_of


def check_lazy_dispatch(n=24, device=0, pairs=(("PERLND", "cbp_fullchain"), ("IMPLND", "i001_test10"))):
    """The drop-in activities called one (segment, activity) at a time by a dispatcher shaped like hsp2.main, over n clones of
    the CBP full-chain PERLND and n of test10's IMPLND 1 with per-segment parameters and forcings: every series left in ts and
    every error vector equals the oracle's run of that segment, while the work was done in one fused batch per operation."""
    import pandas as pd

    from hspsquared_b200 import lazybatch

    cases = CASES.build_cases()
    uci, rows, ddext = {}, [], {}
    _DISPATCH_SERIES.clear()
    si = None
    for op, cname in pairs:  # the unit system is the first case's (siminfo['units'])
        case = cases[cname]
        s = CASES.siminfo_for(case)
        steps = min(24 * 45, s["steps"])
        if si is None:
            si = dict(s, steps=steps, stop=s["start"] + np.timedelta64(steps * int(s["delt"]), "m"), time_unit="us")
        forc = {k: v[:steps] for k, v in CASES.forcings(case).items()}
        for k in range(n):
            seg = "%s%03d" % (op[0], k + 1)
            t = copy.deepcopy(case["tables"])
            f = 1.0 + 0.01 * k
            if op == "PERLND":
                t["PWATER"]["PARAMETERS"]["LZSN"] *= f
                t["PWATER"]["PARAMETERS"]["INFILT"] /= f
            else:
                t["IWATER"]["PARAMETERS"]["RETSC"] *= f
            t["SNOW"]["PARAMETERS"]["SNOWCF"] *= f
            uci[(op, "GENERAL", seg)] = {"ACTIVITY": dict(t["ACTIVITY"])}
            for act, tab in t.items():
                if act != "ACTIVITY":
                    uci[(op, act, seg)] = tab
            rows.append((op, seg, int(s["delt"])))
            ddext[(op, seg)] = (op, seg)
            _DISPATCH_SERIES[(op, seg)] = {kk: (v * f if kk == "PREC" else v.copy()) for kk, v in forc.items()}
    opseq = pd.DataFrame(rows, columns=["OPERATION", "SEGMENT", "INDELT_minutes"])
    activities.DEVICE = device
    registry = {op: {a: fn for a, fn in acts.items() if a not in ("PQUAL", "IQUAL")} for op, acts in activities.ACTIVITIES.items()}
    lazybatch.reset()
    got, got_err = _dispatch_like_main(None, {"opseq": opseq, "uci": copy.deepcopy(uci), "ddext_sources": ddext,
                                              "siminfo": dict(si)}, registry)
    stats = dict(lazybatch.STATS)
    assert stats["windows"] == 2 and stats["segments"] == 2 * n and stats["fallback"] == 0 and stats["launches"] == 2, stats
    for (op, seg), ts in got.items():
        tables = {"ACTIVITY": uci[(op, "GENERAL", seg)]["ACTIVITY"]}
        tables.update({act: copy.deepcopy(tab) for (o, act, s2), tab in uci.items() if o == op and s2 == seg and act != "GENERAL"})
        ref = dict(_DISPATCH_SERIES[(op, seg)])
        ref_err = CASES.run_segment(op, tables, ref, dict(si), ORACLE_ACTS)
        assert set(ref) <= set(ts), sorted(set(ref) - set(ts))
        bad = compare(ref, ts, exact=GPU_EXACT)
        assert not bad, (op, seg, bad[:6])
        for a, e in ref_err.items():
            assert np.array_equal(got_err[(op, seg, a)], e), (op, seg, a)
    return stats


def check_streaming_sink(tmpdir, n=37, days=9, chunk=50, device=0, pinned=False):
    """sink.stream_batch (chunks leave the device while the next one computes, a writer thread puts them into the memory-mapped
    store) against results.save_batch on the whole-run array: every (segment, activity) frame read back from the store equals
    the frame save_timeseries would hand IOManager.write_ts -- same columns (SAVE flags), float32, same values."""
    import pandas as pd

    from hspsquared_b200.results import save_batch
    from hspsquared_b200.sink import NpyResults, stream_batch

    start = np.datetime64("1976-02-01T00:00:00")
    cal = Calendar(start, start + np.timedelta64(days, "D"), 60, "us")
    ens = synth.Ensemble("PERLND", n, seed=8, sections=("SNOW", "PWATER"))
    names = list(synth.PERLND_FORCINGS)
    f = synth.forcing_chunk(ens, cal, 0, cal.steps, names=names)
    tindex = pd.date_range(pd.Timestamp(str(start)), periods=cal.steps + 1, freq="60min")[1:]
    si = {"tindex": tindex}
    segs = ["P%03d" % (i + 1) for i in range(n)]
    tabs = {"SNOW": {k: 1 for k in test10.SAVE_DEFAULT["SNOW"]}, "PWATER": {k: 1 for k in test10.SAVE_DEFAULT["PWATER"]}}
    tabs["PWATER"]["PET"] = 0  # produced, present in the SAVE table, switched off: must not be stored (io.py:70-72)
    order = np.argsort(ens.u_airt, kind="stable")  # a device order that differs from the caller's
    with LandBatch(ens.store(cal), cal, names, "all", device=device, order=order) as b:
        whole = b.run(f, chunk=chunk)
        saved = b.save
    ref = FrameSink()
    save_batch(ref, si, "PERLND", segs, tabs, saved, whole.astype(np.float32))
    sink = NpyResults(os.path.join(str(tmpdir), "results"), mode="w")
    with LandBatch(ens.store(cal), cal, names, "all", device=device, order=order) as b:
        fdev = np.ascontiguousarray(f[:, b._fperm, :][:, :, b.order])
        stream_batch(sink, b, si, "PERLND", segs, tabs, lambda t0, t1: np.ascontiguousarray(fdev[t0:t1]), chunk, pinned=pinned)
    sink.close()
    store = NpyResults(os.path.join(str(tmpdir), "results"))
    assert len(ref.calls) == 2 * n
    for (op, seg, act), (df, save_columns) in ref.calls.items():
        want = df[[c for c in df.columns if c in save_columns]]  # IOManager.write_ts drops the unsaved columns (io.py:70-72)
        got = store.read_ts("RESULT", op, seg, act)
        assert list(got.columns) == list(want.columns), (seg, act)
        assert got.index.equals(want.index)
        assert got.to_numpy().dtype == np.float32
        assert np.array_equal(got.to_numpy(), want.to_numpy(), equal_nan=True), (seg, act)
    assert "PET" not in store.read_ts("RESULT", "PERLND", segs[0], "PWATER").columns


def check_grouped_batch(n=300, device=0, layout="plain", jit=None):
    """engine.GroupedLandBatch (one batch and one kernel per option word, launched together over column ranges of one series
    buffer, hsp2g_run_device_group) against the single mixed batch that tests the options at run time: every segment of a
    full-chain ensemble with per-segment option flags comes out bit for bit the same -- through run() and through the
    device-resident group launch --, state and error counts included."""
    import ctypes as C

    from hspsquared_b200.engine import GroupedLandBatch, tile_series, untile_series

    cal = Calendar("1976-02-25", "1976-03-08", 60, "us")
    steps = cal.steps
    ens = synth.Ensemble("PERLND", n, seed=77, base="cbp", options=("RTOPFG", "UZFG", "SNOPFG", "ICEFG", "TSOPFG"))
    names = list(synth.FULLCHAIN_FORCINGS)
    f = synth.forcing_chunk(ens, cal, 0, steps, names=names)
    save = list(cbp10()) + ["SNOW_PACKF", "PSTEMP_SLTMP"]
    store = ens.store(cal)
    with LandBatch(store, cal, names, save, device=device, jit=False) as b:
        ref = b.run(f, chunk=100)
        ref_state, ref_err = b.get_state(), b.errors()
    lib = _lib.load()
    with GroupedLandBatch(store, cal, names, save, device=device, layout=layout, jit=jit, keys=(ens.u_airt,)) as g:
        assert len(g.groups) == len(np.unique(store.flags)) > 8
        got = g.run(f, chunk=100)
        assert np.array_equal(got, ref, equal_nan=True)
        assert np.array_equal(g.get_state(), ref_state, equal_nan=True)
        e = g.errors()
        assert all(np.array_equal(e[k], ref_err[k]) for k in ref_err)
        if jit:
            assert all(",jit>" in k and "options=0x" in k for k in g.kernel_names()), g.kernel_names()[:3]
    # device-resident: one buffer of `width` columns, the groups launched together
    with GroupedLandBatch(store, cal, names, save, device=device, layout=layout, jit=jit, keys=(ens.u_airt,)) as g:
        W = g.width
        fdev = np.zeros((steps, g.nf, W))
        frows = [names.index(r) for r in g.forcing_rows]
        for c0, idx, b in zip(g.first_column, g.members, g.groups):
            cols = np.concatenate([idx, np.repeat(idx[-1:], b.npad - len(idx))])
            fdev[:, :, c0:c0 + b.npad] = f[:, frows, :][:, :, cols]
        if layout == "tiled":
            fbuf = np.concatenate([tile_series(fdev[:, :, c0:c0 + b.npad]) for c0, b in zip(g.first_column, g.groups)], axis=0)
        else:
            fbuf = np.ascontiguousarray(fdev)
        obytes = steps * g.ns * W * 8
        d_f, d_o = lib.hsp2g_device_alloc(device, fbuf.nbytes), lib.hsp2g_device_alloc(device, obytes)
        assert d_f and d_o
        try:
            _lib.check(lib.hsp2g_memcpy_h2d(device, d_f, fbuf.ctypes.data, fbuf.nbytes))
            for t0 in range(0, steps, 96):  # chunked: the state is carried inside every group
                m = min(96, steps - t0)
                if layout == "plain":
                    g.run_device(t0, m, d_f + t0 * g.nf * W * 8, d_o + t0 * g.ns * W * 8)
                else:  # tile-major buffers hold one chunk: stage the chunk
                    fb = np.concatenate([tile_series(fdev[t0:t0 + m, :, c0:c0 + b.npad]) for c0, b in zip(g.first_column, g.groups)], axis=0)
                    _lib.check(lib.hsp2g_memcpy_h2d(device, d_f, fb.ctypes.data, fb.nbytes))
                    g.run_device(t0, m, d_f, d_o + t0 * g.ns * W * 8)
                g.sync()
            raw = np.empty(steps * g.ns * W)
            _lib.check(lib.hsp2g_memcpy_d2h(device, raw.ctypes.data, d_o, obytes))
        finally:
            lib.hsp2g_device_free(device, C.c_void_p(d_f))
            lib.hsp2g_device_free(device, C.c_void_p(d_o))
        if layout == "plain":
            dev = raw.reshape(steps, g.ns, W)
        else:
            dev = np.empty((steps, g.ns, W))
            for t0 in range(0, steps, 96):
                m = min(96, steps - t0)
                blk = raw[t0 * g.ns * W:(t0 + m) * g.ns * W].reshape(W // 32, m, g.ns, 32)
                dev[t0:t0 + m] = untile_series(blk, W)
        srows = [g.save_rows.index(s_) for s_ in g.save]
        assert np.array_equal(dev[:, srows, :][:, :, g.column_of], ref, equal_nan=True)
        assert np.array_equal(g.get_state(), ref_state, equal_nan=True)
