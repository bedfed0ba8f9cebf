#!/usr/bin/env python
"""Counts of the TMA / mbarrier / store instructions in the SASS of the shipped library and of the headline kernels built per
batch: the tracked evidence that the forcing ring is fed by the TMA engine (UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor,
SYNCS.* = mbarrier) and the saved series leave through evict-first 64-bit stores.   python tools/sass_grep.py > profiles/sass_grep.txt"""
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("UBLKCP", "UTMALDG", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "LDGSTS", "STG.E.EF", "LDS", "DFMA", "MUFU.RCP64H")


def count(path):
    s = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    return " ".join("%s=%d" % (k, len(re.findall(r"\b" + re.escape(k), s))) for k in KEYS)


def main():
    print("# cuobjdump -sass counts; regenerate with: python tools/sass_grep.py > profiles/sass_grep.txt")
    print("hspsquared_b200/libhsp2g.so (the library's general kernels, all instantiations)\n    " +
          count(os.path.join(ROOT, "hspsquared_b200", "libhsp2g.so")))
    for f in sorted(glob.glob(os.path.join(ROOT, "hspsquared_b200", "_kcache", "*.json"))):
        m = json.load(open(f))
        n = m.get("name", "")
        if m.get("kind") == "cuda/sm_100a" and "SNOW|PWATER,hourly,forc=6:0x7e" in n and "save=24" in n and "options=0x330000d" in n \
                and "met=shared" not in n:
            print("%s   (registers %s, spill %s B)\n    %s" % (n, m.get("registers"), m.get("spill_bytes"), count(f.replace(".json", ".cubin"))))


if __name__ == "__main__":
    sys.exit(main())
