"""WDM decoder (hspsquared_b200/wdm.py) against the reference's decoder on every dataset of the reference's WDM files, against
the committed hashes of that decode, and -- where the reference is absent -- on a synthetic file."""
import hashlib
import json
import os

import numpy as np
import pytest

from hspsquared_b200.wdm import WdmFile
from tests.wdm_writer import write_wdm

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (build container only)")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@needs_ref
def test_every_dataset_matches_committed_reference_decode(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "wdm_ref.json")))
    total = 0
    for rel, sets in gold.items():
        w = WdmFile(os.path.join(REF, rel))
        assert [str(d) for d in w.dsns if len(w.series(d))] == sorted(sets, key=int)
        for d, g in sets.items():
            s = w.series(int(d))
            assert len(s) == g["n"]
            assert _sha(s.values) == g["values"], (rel, d)
            assert _sha(s.times.astype(np.int64)) == g["times"], (rel, d)
            total += len(s)
    assert total > 8e7  # 82.7 M values in 246 datasets


@needs_ref
def test_against_reference_walkers_directly():
    from oracle import ref_harness as H

    ref = H.read_wdm(H.TEST10_WDM)
    w = WdmFile(H.TEST10_WDM)
    for d, s in ref.items():
        m = w.series(d)
        assert np.array_equal(m.values, s.to_numpy()) and np.array_equal(m.times, s.index.values.astype("datetime64[s]"))
        step = m.step
        assert step is not None and step == np.timedelta64(int(s.index.freq.nanos // 10 ** 9), "s")
    a = w.attributes(39)
    assert a["TCODE"] == 3 and a["TSSTEP"] == 1 and a["TSTYPE"] == "PREC"


def test_synthetic_file_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    t0 = np.datetime64("1990-01-01T00:00:00", "s")
    hourly = rng.random(24 * 31 + 24 * 28).astype(np.float32)  # two monthly groups of hourly data
    jan, feb = hourly[:24 * 31], hourly[24 * 31:]
    blocks_jan = [(jan[:300], False), (np.zeros(144, np.float32), True), (jan[444:], False)]
    blocks_feb = [(feb[:500], False), (feb[500:], False)]  # the second block starts a new record
    daily = rng.random(365).astype(np.float32)
    path = str(tmp_path / "t.wdm")
    write_wdm(path, [
        {"dsn": 7, "tcode": 3, "tstep": 1, "tgroup": 5,
         "groups": [(t0, blocks_jan), (np.datetime64("1990-02-01T00:00:00", "s"), blocks_feb)]},
        {"dsn": 12, "tcode": 4, "tstep": 1, "tgroup": 6, "groups": [(t0, [(daily, False)])]},
    ])
    w = WdmFile(path)
    assert w.dsns == [7, 12]
    s = w.series(7)
    want = np.concatenate([jan[:300], np.zeros(144, np.float32), jan[444:], feb]).astype(np.float64)
    assert np.array_equal(s.values, want)
    assert np.array_equal(s.times, t0 + np.arange(len(want)) * np.timedelta64(3600, "s"))
    assert s.step == np.timedelta64(3600, "s")
    d = w.series(12)
    assert np.array_equal(d.values, daily.astype(np.float64)) and d.times[-1] == np.datetime64("1990-12-31T00:00:00")
    with pytest.raises(KeyError):
        w.series(99)
    bad = tmp_path / "bad.wdm"
    np.zeros(1024, dtype=np.int32).tofile(str(bad))
    with pytest.raises(ValueError):
        WdmFile(str(bad))
