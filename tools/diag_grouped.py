"""diagnostic: which groups / series / intervals of the grouped tiled run differ from the mixed batch"""
import sys
import numpy as np
sys.path.insert(0, ".")
from hspsquared_b200 import synth, _lib
from hspsquared_b200.calendar import Calendar
from tests.parity import cbp10
from hspsquared_b200.engine import LandBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cal = Calendar("1976-02-25", "1976-03-08", 60, "us")
steps = cal.steps
ens = synth.Ensemble("PERLND", n, seed=77, base="cbp", options=("RTOPFG", "UZFG", "SNOPFG", "ICEFG", "TSOPFG"))
names = list(synth.FULLCHAIN_FORCINGS)
f = synth.forcing_chunk(ens, cal, 0, steps, names=names)
save = list(cbp10()) + ["SNOW_PACKF", "PSTEMP_SLTMP"]
store = ens.store(cal)
with LandBatch(store, cal, names, save, device=0, jit=False) as b:
    ref = b.run(f, chunk=100)
    rows = b.save
words, inv = np.unique(store.flags, return_inverse=True)
for w in range(len(words)):
    idx = np.flatnonzero(inv == w)
    idx = idx[np.lexsort([ens.u_airt[idx]])]
    sub = store.subset(idx)
    fi = np.ascontiguousarray(f[:, :, idx])
    res = {}
    for layout, jit, chunk in (("tiled", True, 100), ("tiled", True, steps), ("tiled", False, 100), ("plain", True, 100)):
        if True:
            if True:
                with LandBatch(sub, cal, names, save, device=0, layout=layout, jit=jit, order=None) as g:
                    got = g.run(fi, chunk=chunk)
                    kn = g.kernel_name()
                ok = np.array_equal(got, ref[:, :, idx], equal_nan=True)
                res[(layout, jit, chunk)] = ok
                if not ok:
                    bad = ~((got == ref[:, :, idx]) | (np.isnan(got) & np.isnan(ref[:, :, idx])))
                    t, r, s = np.nonzero(bad)
                    print("word %#x n=%d %s jit=%s chunk=%d: %d values differ; first t=%d, rows %s, segs %s..., t range %d-%d  %s" % (
                        words[w], len(idx), layout, jit, chunk, bad.sum(), t.min(), sorted(set(np.array(rows)[r]))[:6],
                        sorted(set(s))[:8], t.min(), t.max(), kn[-60:]), flush=True)
                    k = np.argmin(t)
                    print("   first: t=%d row=%s seg=%d got=%r ref=%r" % (t[k], rows[r[k]], s[k], got[t[k], r[k], s[k]], ref[t[k], r[k], idx[s[k]]]))
    print("word %#x n=%d ok=%s" % (words[w], len(idx), all(res.values())), flush=True)
    nbad = globals().get("nbad", 0) + (0 if all(res.values()) else 1)
    if nbad >= 4:
        break
