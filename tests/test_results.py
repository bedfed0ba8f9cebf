"""Output path: save_batch hands an IOManager-like sink exactly the float32 frames the reference's
save_timeseries would build from the oracle's ts (utilities.py:292-333), with the float32 cast done in the kernel."""
import numpy as np
import pandas as pd
import pytest

from hspsquared_b200 import synth, test10
from hspsquared_b200.calendar import Calendar
from hspsquared_b200.engine import LandBatch
from hspsquared_b200.results import save_batch
from oracle import cases as CASES

from . import parity


@pytest.fixture(scope="module", autouse=True)
def _use_emu(emu_library):
    yield


class Sink:
    """Records write_ts calls (the reference's SupportsWriteTS surface as IOManager exposes it, hsp2io/io.py:58-66)."""

    def __init__(self):
        self.calls = {}

    def write_ts(self, data_frame, save_columns, category, operation=None, segment=None, activity=None, compress=True,
                 outstep=2):
        self.calls[(operation, segment, activity)] = (data_frame.copy(), list(save_columns), category, outstep)


def reference_frame(ts, savedict, tindex):
    """save_timeseries' frame for a land activity (utilities.py:308-313), restated for the test."""
    df = pd.DataFrame(index=tindex)
    for y in (savedict.keys() & set(ts.keys())):
        df[y] = ts[y]
    return df.astype(np.float32).sort_index(axis="columns")


def test_save_batch_frames_equal_reference_frames():
    cal = Calendar("1976-01-01", "1976-02-15", 60, "us")
    steps = cal.steps
    tindex = pd.date_range("1976-01-01", "1976-02-15", freq="60min")[:-1]
    assert len(tindex) == steps
    n = 40
    ens = synth.Ensemble("PERLND", n, seed=11, sections=("SNOW", "PWATER"))
    save_tables = {"SNOW": {k: 1 for k in test10.SAVE_DEFAULT["SNOW"]}, "PWATER": {k: 1 for k in test10.SAVE_DEFAULT["PWATER"]}}
    save_tables["PWATER"]["TAET"] = 0  # present in SAVE with flag 0: a column of the frame, not of save_columns
    save = ["SNOW_" + k for k in save_tables["SNOW"]] + ["PWATER_" + k for k in save_tables["PWATER"]]
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), save) as b:
        b.set_output_dtype(np.float32)
        out = b.run(synth.forcing_chunk(ens, cal, 0, steps), chunk=400)
        rows = b.save
    assert out.dtype == np.float32
    sink = Sink()
    names = ["P%03d" % (i + 1) for i in range(n)]
    save_batch(sink, {"tindex": tindex}, "PERLND", names, save_tables, rows, out, outstep=3)
    assert len(sink.calls) == 2 * n
    for i in (0, 17, 39):
        si = {"start": cal.start, "stop": cal.stop, "delt": 60, "units": 1, "steps": steps, "time_unit": "us"}
        ts = synth.forcing_of(ens, cal, i, 0, steps)
        CASES.run_segment("PERLND", ens.tables_of(i), ts, si, parity.ORACLE_ACTS)
        for act in ("SNOW", "PWATER"):
            df, cols, cat, outstep = sink.calls[("PERLND", names[i], act)]
            ref = reference_frame(ts, save_tables[act], tindex)
            assert list(df.columns) == list(ref.columns)
            assert df.index.equals(ref.index) and outstep == 3
            assert np.array_equal(df.to_numpy(), ref.to_numpy(), equal_nan=True)
            assert df.to_numpy().dtype == np.float32
            assert cols == [k for k, v in save_tables[act].items() if v]
    assert "TAET" in sink.calls[("PERLND", "P001", "PWATER")][0].columns
    assert "TAET" not in sink.calls[("PERLND", "P001", "PWATER")][1]


def test_save_qual_batch_frames_equal_reference_frames():
    """IQUAL rows -> one frame per segment with the reference's substring-matched column set (utilities.py:294-301)."""
    from hspsquared_b200.engine import QualBatch
    from hspsquared_b200.results import save_qual_batch

    case = CASES.build_cases()["i_iqual_variants"]
    ref, _ = parity.run_oracle(case)
    si = CASES.siminfo_for(case)
    cal = Calendar(si["start"], si["stop"], si["delt"], "us")
    steps = case["steps"]
    tindex = pd.date_range("1976-01-01", periods=steps + 1, freq="60min")[1:]
    n = 3
    series = {k: ref[k] for k in ("SURO", "SOSLD", "PREC", "SLIQSP")}
    with QualBatch("IMPLND", [case["tables"]["IQUAL"]] * n, cal, list(series), "all") as q:
        out = q.run(q.gather(series))
        savedict = {k: 0 for k in ("SQO", "SOQSP", "SOQS", "SOQO", "SOQOC", "SOQUAL", "SOQC", "IQADDR", "IQADWT", "IQADEP")}
        savedict["SOQUAL"] = 1
        sink = Sink()
        save_qual_batch(sink, {"tindex": tindex}, "IMPLND", ["I001", "I002", "I003"], savedict, q, out)
    assert len(sink.calls) == n
    # the reference's frame, restated: every ts key that contains '_<key>'
    want = pd.DataFrame(index=tindex)
    for y in savedict:
        for z in ref:
            if "_" + y in z:
                want[z] = ref[z]
    want = want.astype(np.float32).sort_index(axis="columns")
    assert len(want.columns) == 3 * 10 and "IQUAL2_SOQOC" in want.columns and "IQUAL1_SLIQO" not in want.columns
    for seg in ("I001", "I003"):
        df, cols, cat, outstep = sink.calls[("IMPLND", seg, "IQUAL")]
        assert list(df.columns) == list(want.columns) and cols == ["SOQUAL"]
        assert np.array_equal(df.to_numpy(), want.to_numpy(), equal_nan=True) and df.to_numpy().dtype == np.float32


def test_period_aggregation_equals_pandas():
    parity.check_period_aggregation()


def test_daily_bins_are_anchored_at_the_first_label():
    """IOManager.write_ts resamples with ('D', origin='start') (hsp2io/io.py:76-78): 'D' is a 24-hour Tick in the pandas the
    reference pins, so the bins start at the frame's first label.  A 48-hour hourly run from midnight therefore has 2 periods of
    24 values labelled 01:00 -- not 23 / 24 / 1 values on calendar days."""
    import pandas as pd

    from hspsquared_b200.results import period_bins

    tindex = pd.date_range("1976-03-01", periods=49, freq="60min")[1:]  # interval-END labels, main.py:94
    bins, labels = period_bins(tindex, 3)
    assert np.array_equal(np.bincount(bins), [24, 24])
    assert list(labels) == [pd.Timestamp("1976-03-01 01:00"), pd.Timestamp("1976-03-02 01:00")]
    want = pd.Series(np.arange(48.0), index=tindex).resample("24h", origin="start").sum()
    assert want.index.equals(labels) and list(want) == [float(sum(range(24))), float(sum(range(24, 48)))]
    # a run that starts mid-day, 30-minute step: still whole 24-hour periods from the first label
    t2 = pd.date_range("1976-03-01 14:00", periods=48 * 3 + 1, freq="30min")[1:]
    b2, l2 = period_bins(t2, 3)
    assert np.array_equal(np.bincount(b2), [48, 48, 48]) and l2[0] == pd.Timestamp("1976-03-01 14:30")
    w2 = pd.Series(np.ones(len(t2)), index=t2).resample("24h", origin="start").size()
    assert w2.index.equals(l2)
    # months and years are calendar periods (origin= has no effect on calendar offsets in any pandas)
    t3 = pd.date_range("1976-01-01", periods=24 * 70 + 1, freq="60min")[1:]
    b3, l3 = period_bins(t3, 4)
    assert np.array_equal(np.bincount(b3), [24 * 31 - 1, 24 * 29, 24 * 10 + 1]) and l3[0] == pd.Timestamp("1976-01-31")
