#!/usr/bin/env python
"""Diagnostic: strictly relative error |gpu - oracle| / |oracle| (no scale floor) of every saved series of every parity
case, to see how much slack the test metric (tests/parity.py rel_err: denominator floored at 1e-3 of the series' scale)
actually uses.  Prints the worst offenders."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # run as: python tests/strict_parity.py
sys.path.insert(0, ROOT)
from oracle import cases as CASES  # noqa: E402
from tests import parity  # noqa: E402


def main():
    rows = []
    for name, case in sorted(CASES.build_cases().items()):
        ref, _ = parity.run_oracle(case)
        got, _, kname = parity.run_fused(case, replicate=2)
        for k, g in got.items():
            r = ref[k][:len(g)]
            g0 = g[:, 0]
            ok = ~(np.isnan(r) | (np.abs(r) > 1e29))
            d = np.abs(g0[ok] - r[ok])
            nz = np.abs(r[ok]) > 0
            rel = d[nz] / np.abs(r[ok][nz])
            strict = float(rel.max()) if nz.any() else 0.0
            if nz.any() and strict > 1e-9:
                scale = float(np.abs(r[ok]).max())
                bad = rel > 1e-9
                print("%-18s %-7s strict %.2e: %d samples > 1e-9, largest |oracle| among them %.3e = %.1e of the series scale, "
                      "largest |diff| %.1e" % (name, k, strict, int(bad.sum()), float(np.abs(r[ok][nz])[bad].max()),
                                               float(np.abs(r[ok][nz])[bad].max()) / scale, float(d[nz][bad].max())))
            absz = float(d[~nz].max()) if (~nz).any() else 0.0  # error where the oracle is exactly zero
            rows.append((strict, absz, name, k, float(np.abs(r[ok]).max()) if ok.any() else 0.0))
    rows.sort(reverse=True)
    print("worst strict relative errors (rel, abs-at-zero, case, series, scale):")
    for r in rows[:25]:
        print("%.3e %.3e %-20s %-8s %.3e" % r)
    print("worst absolute error where the oracle is exactly 0:", max(r[1] for r in rows))
    print("series with strict rel > 1e-9:", sum(1 for r in rows if r[0] > 1e-9), "of", len(rows))


if __name__ == "__main__":
    main()
