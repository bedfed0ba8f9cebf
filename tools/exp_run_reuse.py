#!/usr/bin/env python
"""LandBatch.run() with ordinary numpy arrays, headline shape (100 k segments x 438 intervals): a new output array per call
against the same arrays handed in again (registrations kept, engine._PageLocked)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hspsquared_b200 import synth  # noqa: E402
from hspsquared_b200.calendar import Calendar  # noqa: E402
from hspsquared_b200.engine import LandBatch  # noqa: E402


def main():
    n, Te = 100_000, 438
    cal = Calendar(bench.RUN_START, bench.RUN_STOP, 60, "ns")
    ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
    f = synth.forcing_chunk(ens, cal, 0, Te, names=list(synth.PERLND_FORCINGS))
    res = {"segments": n, "intervals": Te, "bytes_in": f.nbytes}
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), bench.default_save()) as b:
        b.set_output_dtype(np.float32)
        t = time.perf_counter()
        o = b.run(f, chunk=Te)
        res["first_call_s"] = time.perf_counter() - t
        res["bytes_out"] = o.nbytes
        new, again = [], []
        for _ in range(4):
            t = time.perf_counter()
            o2 = b.run(f, chunk=Te)
            new.append(time.perf_counter() - t)
            del o2
        for _ in range(6):
            t = time.perf_counter()
            b.run(f, chunk=Te, out=o)
            again.append(time.perf_counter() - t)
        res["new_output_array_s"] = new
        res["same_arrays_again_s"] = again
        res["same_arrays_again_value"] = n * Te / min(again)
        res["checksum"] = float(np.nansum(o[-1, 13:], dtype=np.float64))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
