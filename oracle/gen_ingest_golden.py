#!/usr/bin/env python
"""Fixtures for the forcing-ingest path, generated from the UNMODIFIED reference in the build container.

    python -m oracle.gen_ingest_golden

Writes
  tests/golden/ingest_ref.json  sha256 (and length, NaN count, sum) of the outputs of the reference's utilities.transform
                                (utilities.py:83-133) for seeded synthetic sources over a grid of (run step, source step,
                                member, transformation); tests/test_ingest.py regenerates the inputs from the seed
  tests/golden/wdm_ref.json     per dataset of the reference's WDM files: length, first / last time stamp and sha256 of the
                                time stamps and values decoded by the reference's walkers (hsp2tools/readWDM.py:166-319)
"""
import glob
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
RUN = ("1976-01-03", "1976-02-20")
SOURCES = (("2h", 7200, 3), ("h", 3600, 3), ("D", 86400, 4), ("15min", 900, 2), ("30min", 1800, 2), ("3h", 10800, 3))
HOWS = ("SAME", "DIV", "", "MEAN", "SUM", "MAX", "MIN", "LAST", "INTERPOLATE")


def source_values(freq, step):
    """Seeded synthetic source covering 1976-01-01 + 70 days (the test regenerates exactly this)."""
    n = int(70 * 86400 / step)
    rng = np.random.default_rng(abs(hash_str(freq)) % (2 ** 31))
    return np.datetime64("1976-01-01", "s") + np.arange(n) * np.timedelta64(step, "s"), rng.random(n) * 3.0


def hash_str(s):
    return int(hashlib.sha256(s.encode()).hexdigest()[:8], 16)


def gen_transform():
    import pandas as pd

    out = {}
    for delt in (60, 30, 120):
        si = H.make_siminfo(RUN[0], RUN[1], delt)
        for freq, step, _tcode in SOURCES:
            t, v = source_values(freq, step)
            idx = pd.DatetimeIndex(t)
            idx.freq = pd.tseries.frequencies.to_offset(freq)
            for name in ("PREC", "AIRTMP"):
                for how in HOWS:
                    s = pd.Series(v.copy(), index=idx)
                    try:
                        r = H.transform(s, name, how, si)
                    except Exception:  # the reference cannot evaluate this combination under this pandas: not pinned
                        continue
                    out["%d|%s|%s|%s" % (delt, freq, name, how)] = {
                        "n": len(r), "nan": int(np.isnan(r).sum()), "sum": float(np.nansum(r)),
                        "sha256": hashlib.sha256(np.ascontiguousarray(r, dtype=np.float64).tobytes()).hexdigest()}
    with open(os.path.join(GOLD, "ingest_ref.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("transform cases:", len(out))


def gen_wdm():
    out = {}
    for p in sorted(glob.glob(os.path.join(H.REF_ROOT, "tests", "**", "*.wdm"), recursive=True)):
        rel = os.path.relpath(p, H.REF_ROOT)
        ref = H.read_wdm(p)
        out[rel] = {str(d): {"n": len(s), "first": str(s.index[0]), "last": str(s.index[-1]),
                             "freq": str(s.index.freq.freqstr) if s.index.freq is not None else None,
                             "values": hashlib.sha256(np.ascontiguousarray(s.to_numpy()).tobytes()).hexdigest(),
                             "times": hashlib.sha256(np.ascontiguousarray(
                                 s.index.values.astype("datetime64[s]").astype(np.int64)).tobytes()).hexdigest()}
                    for d, s in ref.items()}
        print(rel, len(ref))
    with open(os.path.join(GOLD, "wdm_ref.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    gen_transform()
    gen_wdm()
