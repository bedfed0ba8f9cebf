#!/usr/bin/env python
"""Where the time of LandBatch.run() with ordinary numpy arrays goes (headline shape: 100 k segments x 438 intervals)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hspsquared_b200 import _lib, synth  # noqa: E402
from hspsquared_b200.calendar import Calendar  # noqa: E402
from hspsquared_b200.engine import LandBatch, _prefault  # noqa: E402


def main():
    n, Te = 100_000, 438
    cal = Calendar(bench.RUN_START, bench.RUN_STOP, 60, "ns")
    ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
    f = synth.forcing_chunk(ens, cal, 0, Te, names=list(synth.PERLND_FORCINGS))
    out = {}
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), bench.default_save()) as b:
        lib = b.lib
        b.set_output_dtype(np.float32)
        b.run(f, chunk=Te)
        for rep in range(2):
            T = {}
            t = time.perf_counter()
            dev = np.empty((Te, b.ns, n), dtype=np.float32)
            T["empty"] = time.perf_counter() - t
            t = time.perf_counter()
            _prefault(dev)
            T["prefault"] = time.perf_counter() - t
            t = time.perf_counter()
            assert lib.hsp2g_host_register(f.ctypes.data, f.nbytes) == 0
            T["register_in"] = time.perf_counter() - t
            t = time.perf_counter()
            assert lib.hsp2g_host_register(dev.ctypes.data, dev.nbytes) == 0
            T["register_out"] = time.perf_counter() - t
            t = time.perf_counter()
            b._kernel()
            _lib.check(lib.hsp2g_run_host_ld(b.h, 0, Te, f.ctypes.data, n, dev.ctypes.data, n))
            T["enqueue"] = time.perf_counter() - t
            t = time.perf_counter()
            b.sync()
            T["sync"] = time.perf_counter() - t
            t = time.perf_counter()
            lib.hsp2g_host_unregister(f.ctypes.data)
            T["unregister_in"] = time.perf_counter() - t
            t = time.perf_counter()
            lib.hsp2g_host_unregister(dev.ctypes.data)
            T["unregister_out"] = time.perf_counter() - t
            t = time.perf_counter()
            o = b.run(f, chunk=Te)
            T["run_new_array"] = time.perf_counter() - t
            t = time.perf_counter()
            b.run(f, chunk=Te, out=o)
            T["run_reused_array"] = time.perf_counter() - t
            del dev, o
            out["rep%d" % rep] = {k: round(v, 4) for k, v in T.items()}
    out["bytes_in"], out["bytes_out"] = f.nbytes, Te * 24 * n * 4
    print(json.dumps(out))


if __name__ == "__main__":
    main()
