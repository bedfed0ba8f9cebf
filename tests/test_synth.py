"""Synthetic ensembles: the vectorised segment store equals the per-segment table path, and a jittered
ensemble run through the (test-double) library matches the CPU oracle segment by segment."""

import numpy as np
import pytest

from hspsquared_b200 import synth
from hspsquared_b200.calendar import Calendar
from hspsquared_b200.engine import LandBatch
from hspsquared_b200.segstore import SegmentStore
from oracle import cases as CASES

from . import parity


@pytest.fixture(scope="module", autouse=True)
def _use_emu(emu_library):
    yield


@pytest.mark.parametrize("op,sections", [("PERLND", ("SNOW", "PWATER")), ("IMPLND", ("SNOW", "IWATER"))])
def test_vectorised_store_equals_table_path(op, sections):
    cal = Calendar("1976-01-01", "1976-03-01", 60, "us")
    ens = synth.Ensemble(op, 37, seed=5, sections=sections)
    a = ens.store(cal)
    b = SegmentStore.from_tables(op, [ens.tables_of(i) for i in range(ens.n)], cal)
    assert np.array_equal(a.par, b.par, equal_nan=True)
    assert np.array_equal(a.state, b.state, equal_nan=True)
    assert np.array_equal(a.flags, b.flags)
    assert set(a.monthly) == set(b.monthly)
    for k in a.monthly:
        assert np.array_equal(a.monthly[k], b.monthly[k])


def test_met_index_drops_feb29():
    cal = Calendar("1976-01-01", "1979-01-01", 60, "us")
    idx = synth.met_index(cal)
    assert len(idx) == (366 + 365 + 365) * 24
    assert idx[:8784].tolist() == list(range(8784))
    y77 = idx[8784:8784 + 8760]
    assert y77[59 * 24 - 1] == 59 * 24 - 1 and y77[59 * 24] == 60 * 24  # Feb 28 23:00 -> Mar 1 00:00
    assert y77[-1] == 8783


def test_ensemble_matches_oracle_per_segment():
    cal = Calendar("1976-01-01", "1976-04-10", 60, "us")
    steps = cal.steps
    ens = synth.Ensemble("PERLND", 130, seed=1234, sections=("SNOW", "PWATER"))
    store = ens.store(cal)
    names = list(synth.PERLND_FORCINGS)
    with LandBatch(store, cal, names, "all") as b:
        out = b.run(synth.forcing_chunk(ens, cal, 0, steps), chunk=500)
        save = b.save
    for i in (0, 63, 129):
        si = {"start": cal.start, "stop": cal.stop, "delt": 60, "units": 1, "steps": steps, "time_unit": "us"}
        ts = synth.forcing_of(ens, cal, i, 0, steps)
        CASES.run_segment("PERLND", ens.tables_of(i), ts, si, parity.ORACLE_ACTS)
        for r, full in enumerate(save):
            k = full.split("_", 1)[1]
            assert np.array_equal(out[:, r, i], ts[k], equal_nan=True), (i, k)


def test_categorical_fullchain_ensemble_matches_oracle_and_sorting_is_invisible():
    """BASELINE.json configs[3]: full PERLND chain on the CBP base with option flags drawn per segment.  Every segment
    equals the oracle run on its own tables bit for bit, and divergence-aware ordering (SegmentStore.sort_by_flags)
    only permutes columns."""
    cal = Calendar("1976-02-20", "1976-04-05", 60, "us")
    steps = cal.steps
    opts = ("RTOPFG", "UZFG", "SNOPFG", "ICEFG", "TSOPFG")
    ens = synth.Ensemble("PERLND", 100, seed=99, base="cbp", options=opts)
    assert len(np.unique(ens.choice.T, axis=0)) > 20
    names = list(synth.FULLCHAIN_FORCINGS)
    f = synth.forcing_chunk(ens, cal, 0, steps, names=names)
    store = ens.store(cal)
    assert len(np.unique(store.flags)) > 20
    with LandBatch(store, cal, names, "all") as b:
        out = b.run(f, chunk=300)
        save = b.save
        errs = b.errors()
    assert len(save) == 84  # AIRTMP + 22 SNOW + 31 PWATER + 5 SEDMNT + 4 PSTEMP + 21 PWTGAS
    for i in (0, 17, 42, 77, 99):
        si = {"start": cal.start, "stop": cal.stop, "delt": 60, "units": 1, "steps": steps, "time_unit": "us"}
        ts = synth.forcing_of(ens, cal, i, 0, steps)
        e = CASES.run_segment("PERLND", ens.tables_of(i), ts, si, parity.ORACLE_ACTS)
        for r, full in enumerate(save):
            k = full.split("_", 1)[1]
            assert np.array_equal(out[:, r, i], ts[k], equal_nan=True), (i, k)
        assert np.array_equal(e["PWATER"], np.array([errs["PWATER%d" % j][i] for j in range(10)]))
    sorted_store = ens.store(cal)
    perm = sorted_store.sort_by_flags()
    assert np.all(np.diff(sorted_store.flags.astype(np.int64)) >= 0)
    with LandBatch(sorted_store, cal, names, "all") as b:
        out2 = b.run(np.ascontiguousarray(f[:, :, perm]), chunk=300)
    assert np.array_equal(out2, out[:, :, perm], equal_nan=True)


def test_landbatch_order_is_invisible_to_the_caller():
    """LandBatch(order=...) reorders segments on the device (divergence-aware ordering with caller keys) but run(),
    errors() and get_state() keep the caller's order, bit for bit."""
    cal = Calendar("1976-03-01", "1976-03-25", 60, "us")
    ens = synth.Ensemble("PERLND", 90, seed=5, base="cbp", options=("RTOPFG", "UZFG", "ICEFG"))
    names = list(synth.FULLCHAIN_FORCINGS)
    f = synth.forcing_chunk(ens, cal, 0, cal.steps, names=names)
    store = ens.store(cal)
    with LandBatch(store, cal, names, "all") as b:
        ref, ref_e, ref_s = b.run(f, chunk=211), b.errors(), b.get_state()
    perm = store.divergence_order(keys=(np.floor(ens.u_airt * 4), ens.u_prec))
    assert not np.array_equal(perm, np.arange(90))
    flags_sorted = store.flags[perm]
    assert np.all(np.diff(flags_sorted.astype(np.int64)) >= 0)
    for order in ("flags", perm):
        with LandBatch(store, cal, names, "all", order=order) as b:
            out, e, st = b.run(f, chunk=211), b.errors(), b.get_state()
            assert b.order is not None
        assert np.array_equal(out, ref, equal_nan=True)
        assert np.array_equal(st, ref_s, equal_nan=True)
        for k in ref_e:
            assert np.array_equal(e[k], ref_e[k]), k
    assert np.array_equal(store.flags, ens.store(cal).flags)  # the caller's store is untouched


def test_shared_met_mode_equals_expanded_forcings():
    parity.check_shared_met_mode(n=75)
