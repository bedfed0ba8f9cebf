#!/usr/bin/env python
"""What page-locking costs on this host: cudaHostAlloc and cudaHostRegister per GB, and the speed of copying a numpy array into
pinned memory with 1..16 threads -- the numbers behind LandBatch.run()'s handling of ordinary (pageable) numpy arrays."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hspsquared_b200 import _lib  # noqa: E402
from hspsquared_b200.engine import pinned_empty  # noqa: E402


def main():
    lib = _lib.load()
    n = int(lib.hsp2g_device_count(None) if False else 1)
    out = {}
    gb = 2.0
    nel = int(gb * 2**30 / 8)
    for rep in range(2):
        t = time.time()
        p = pinned_empty((nel,))
        out["host_alloc_s_per_gb_%d" % rep] = (time.time() - t) / gb
        t = time.time()
        p[:] = 1.0
        out["first_touch_pinned_s_per_gb_%d" % rep] = (time.time() - t) / gb
        del p
    a = np.ones(nel)
    for rep in range(2):
        t = time.time()
        _lib.check(lib.hsp2g_host_register(a.ctypes.data, a.nbytes))
        out["host_register_s_per_gb_%d" % rep] = (time.time() - t) / gb
        t = time.time()
        _lib.check(lib.hsp2g_host_unregister(a.ctypes.data))
        out["host_unregister_s_per_gb_%d" % rep] = (time.time() - t) / gb
    p = pinned_empty((nel,))
    p[:] = 0.0
    for th in (1, 2, 4, 8, 16):
        cuts = np.linspace(0, nel, th * 4 + 1).astype(np.int64)
        with ThreadPoolExecutor(th) as ex:
            t = time.time()
            list(ex.map(lambda i: np.copyto(p[cuts[i]:cuts[i + 1]], a[cuts[i]:cuts[i + 1]]), range(len(cuts) - 1)))
            out["copy_to_pinned_gbs_%dthreads" % th] = gb * 2**30 / 1e9 / (time.time() - t)
    t = time.time()
    b = np.empty(nel)
    b[:] = 0
    out["pageable_alloc_touch_s_per_gb"] = (time.time() - t) / gb
    out["cpus"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
