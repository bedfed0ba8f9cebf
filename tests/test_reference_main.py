"""The drop-in boundary exercised where the reference itself calls it: the UNMODIFIED `hsp2.hsp2.main.main()` runs test10's
PERLND P001 and IMPLND I001 (SNOW, IWATER, SOLIDS, IQUAL) over an in-memory IO backend (oracle/ref_main.py) -- once with the reference's
own Numba activities and once with `hspsquared_b200.activities.install()` patched into its registry
(configuration.py:45-55) -- and every frame its `save_timeseries` hands to `IOManager.write_ts` must be identical.
Build-container only: needs /root/reference (skipped elsewhere); the CUDA library is replaced by the CPU test double."""
import numpy as np
import pytest

from oracle import ref_harness as H

pytestmark = pytest.mark.skipif(not H.available(), reason="reference tree not present (only in the build container)")


@pytest.fixture(scope="module", autouse=True)
def _use_emu(emu_library):
    yield


@pytest.mark.parametrize("saveall", [False, True])
def test_main_with_dropin_activities_writes_the_same_frames(saveall):
    from hspsquared_b200 import activities
    from oracle import ref_main as R

    uci, wdm = R.test10_land_model("1976-01-01", "1976-03-01")
    ref = R.run_main(uci, wdm, saveall=saveall)
    cfg = H.load()["configuration"]
    stock = {op: dict(acts) for op, acts in cfg.activities.items() if op in ("PERLND", "IMPLND")}
    saved = activities.install(cfg)
    try:
        for op in ("PERLND", "IMPLND"):  # the registry really dispatches to the drop-ins
            for a, fn in activities.ACTIVITIES[op].items():
                assert cfg.activities[op][a] is fn and stock[op][a] is not fn
        got = R.run_main(uci, wdm, saveall=saveall)
    finally:
        activities.uninstall(saved, cfg)
    assert cfg.activities["PERLND"]["PWATER"] is stock["PERLND"]["PWATER"]
    assert set(got) == set(ref) and len(ref) >= 4
    if saveall:
        assert "IQUAL1_SOQUAL" in ref[("IMPLND", "I001", "IQUAL")].columns  # the constituent's series went through too
    for key, df in ref.items():
        g = got[key]
        assert list(g.columns) == list(df.columns), key
        assert g.index.equals(df.index), key
        a, b = g.to_numpy(), df.to_numpy()
        if a.shape[1] == 0:  # IQUAL's SAVE flags are all 0: without saveall IOManager drops every column
            assert key[2] == "IQUAL" and not saveall and b.shape == a.shape
            continue
        assert a.dtype == b.dtype == np.float32
        assert np.array_equal(a, b, equal_nan=True), (key, [c for i, c in enumerate(df.columns)
                                                            if not np.array_equal(a[:, i], b[:, i], equal_nan=True)])


def test_main_in_a_metric_run_writes_the_same_frames():
    """siminfo['units'] = 2 through the unmodified main(): every section of test10's PERLND 1 (SNOW, PWATER, PSTEMP, PWTGAS ...)
    and IMPLND 1 (SNOW, IWATER, SOLIDS, IQUAL) converts where the reference does -- frames identical to the stock Numba run."""
    from hspsquared_b200 import activities, lazybatch
    from oracle import ref_main as R

    uci, wdm = R.test10_land_model("1976-01-01", "1976-03-01", units=2)
    ref = R.run_main(uci, wdm, saveall=True)
    cfg = H.load()["configuration"]
    saved = activities.install(cfg)
    lazybatch.reset()
    try:
        got = R.run_main(uci, wdm, saveall=True)
        stats = dict(lazybatch.STATS)
    finally:
        activities.uninstall(saved, cfg)
    assert set(got) == set(ref)
    for key in (("PERLND", "P001", "PSTEMP"), ("PERLND", "P001", "PWTGAS"), ("IMPLND", "I001", "SOLIDS")):
        assert key in ref and ref[key].shape[1] >= 2, key  # the chain sections' frames are there to be compared
    for key, df in ref.items():
        g = got[key]
        assert list(g.columns) == list(df.columns) and g.index.equals(df.index), key
        assert np.array_equal(g.to_numpy(), df.to_numpy(), equal_nan=True), key
    # one segment per operation: below the lazy-batching threshold, so every call ran the one-section path, each section finding
    # the converted series of the earlier ones in ts (the fused metric chain is test_emu_parity's metric dispatcher test)
    assert stats["fallback"] > 0 and stats["windows"] == 0 and stats["served"] == 0, stats


def test_main_over_many_segments_is_served_from_one_fused_batch():
    """SURVEY 8b "granularity mismatch": the unmodified main() walks 70 PERLND + 70 IMPLND clones (parameters and met
    factors varied per segment) one (segment, activity) at a time; the drop-ins advance each operation's segments in ONE fused
    batch on the first call and serve every later call from it.  Frames identical to the stock Numba run, a handful of kernel
    launches instead of one per (segment, activity)."""
    from hspsquared_b200 import activities, lazybatch
    from oracle import ref_main as R

    uci, wdm = R.test10_land_model("1976-01-01", "1976-02-15", clones=70)
    ref = R.run_main(uci, wdm)
    cfg = H.load()["configuration"]
    saved = activities.install(cfg)
    lazybatch.reset()
    try:
        got = R.run_main(uci, wdm)
        stats = dict(lazybatch.STATS)
    finally:
        activities.uninstall(saved, cfg)
    assert set(got) == set(ref) and len(ref) == 56 * 2 + 14 + 70 * 2  # (segment, activity) frames
    for key, df in ref.items():
        g = got[key]
        assert list(g.columns) == list(df.columns) and g.index.equals(df.index), key
        assert np.array_equal(g.to_numpy(), df.to_numpy(), equal_nan=True), key
    # two windows (one per operation), every land section of every segment served from them, one launch per window and chunk
    assert stats["windows"] == 2 and stats["segments"] == 140, stats
    assert stats["served"] == len(ref) and stats["fallback"] == 0, stats
    assert stats["launches"] == 3, stats  # PERLND with and without SNOW (two activity masks), IMPLND
