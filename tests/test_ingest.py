"""Forcing ingest (hspsquared_b200/ingest.py): `resample` against the reference's `transform` (committed hashes, and the
function itself where the reference tree is present), `MetStore` against the reference's get_timeseries conventions on
test10's EXT SOURCES, and the shared-met kernels fed by a MetStore against the same kernels fed per-segment arrays."""
import hashlib
import json
import os

import numpy as np
import pytest

from hspsquared_b200.calendar import Calendar
from hspsquared_b200.ingest import FLOWTYPE, MetStore, SourceRow, resample, target_name
from hspsquared_b200.wdm import WdmFile
from oracle.gen_ingest_golden import HOWS, RUN, SOURCES, source_values
from tests import parity

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (build container only)")


def test_resample_matches_committed_reference_transform(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "ingest_ref.json")))
    seen = 0
    for delt in (60, 30, 120):
        cal = Calendar(RUN[0], RUN[1], delt, "ns")
        for freq, step, tcode in SOURCES:
            t, v = source_values(freq, step)
            for name in ("PREC", "AIRTMP"):
                for how in HOWS:
                    g = gold.get("%d|%s|%s|%s" % (delt, freq, name, how))
                    if g is None:  # the reference could not evaluate it under the generating pandas
                        continue
                    try:
                        r = resample(t, v, np.timedelta64(step, "s"), name, how, cal, tcode)
                    except NotImplementedError:
                        assert how == "INTERPOLATE"
                        continue
                    assert len(r) == g["n"] and int(np.isnan(r).sum()) == g["nan"], (delt, freq, name, how)
                    assert hashlib.sha256(np.ascontiguousarray(r).tobytes()).hexdigest() == g["sha256"], (delt, freq, name, how)
                    seen += 1
    assert seen > 300


def test_resample_refuses_what_it_cannot_place():
    cal = Calendar("1976-01-03", "1976-01-05", 60, "ns")
    t = np.datetime64("1976-01-04", "s") + np.arange(48) * np.timedelta64(3600, "s")
    with pytest.raises(ValueError):
        resample(t, np.ones(48), np.timedelta64(3600, "s"), "PREC", "SAME", cal, 3)  # source starts after the run
    t2 = np.datetime64("1976-01-01", "s") + np.arange(20) * np.timedelta64(86400, "s")
    with pytest.raises(ValueError):
        resample(t2, np.ones(20), np.timedelta64(86400, "s"), "PREC", "NOPE", cal, 4)
    with pytest.raises(NotImplementedError):
        resample(t2, np.ones(20), np.timedelta64(86400, "s"), "PREC", "ZEROFILL", cal, 4)
    assert target_name("PREC", "1") == "PREC" and target_name("SURLI", "") == "SURLI" and "PETINP" in FLOWTYPE


def test_metstore_layout_and_shared_kernels(emu_library):
    parity.check_metstore_shared()


def test_watershed_wdm_to_frames(emu_library, tmp_path):
    """WDM file -> MetStore -> shared-met full-chain kernels -> save_timeseries frames, against the oracle."""
    sink = parity.check_watershed(tmp_path)
    assert sink.bytes == 48 * 84 * 40 * 24 * 4


@needs_ref
def test_metstore_on_test10_ext_sources_matches_reference():
    """test10's EXT SOURCES (test10.uci:891-914) through WdmFile + MetStore == the arrays the reference's get_timeseries
    conventions give (oracle/ref_harness.test10_forcings: reference WDM walkers, x MFACTOR, reference transform, sum)."""
    from oracle import ref_harness as H

    w = WdmFile(H.TEST10_WDM)
    cal = Calendar("1976-01-01", "1977-01-01", 60, "ns")
    for op in ("PERLND", "IMPLND"):
        want, si = H.test10_forcings(op)
        assert si["steps"] == cal.steps
        names = [k for k in want]
        rows = [SourceRow("TS%03d" % dsn, mf, tran, nm) for dsn, mf, tran, nm in H.TEST10_EXT[op]]
        ms = MetStore.build(cal, names, [rows, rows], lambda k: w.series(int(k[2:])))
        assert ms.n_stations == 1
        full = ms.expand()
        for j, nm in enumerate(names):
            assert np.array_equal(full[:, j, 0], want[nm]), (op, nm)
            assert np.array_equal(full[:, j, 1], want[nm])
    # a scaled rain gage: SAME keeps the factor per segment, and series * factor is the reference's own product
    rows2 = [SourceRow("TS039", 1.25, "SAME", "PREC")]
    ms = MetStore.build(cal, ["PREC"], [rows2], lambda k: w.series(int(k[2:])))
    s = H.read_wdm(H.TEST10_WDM)[39]
    s2 = s * 1.25
    s2.index.freq = s.index.freq
    assert ms.mfactor[0, 0] == 1.25
    assert np.array_equal(ms.expand()[:, 0, 0], H.transform(s2, "PREC", "SAME", H.make_siminfo("1976-01-01", "1977-01-01")))
