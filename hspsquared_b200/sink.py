"""A results store behind HSP2IO's output protocols, and the streaming writer that feeds it chunk by chunk.

The reference ends every (segment, activity) with `save_timeseries` -> `IOManager.write_ts` -> backend.write_ts
(utilities.py:292-333, hsp2io/io.py:58-95, hsp2io/protocols.py:35-43), its only shipped backend being PyTables HDF5
(hsp2io/hdf.py:96-113: one compressed table per `RESULTS/<OPERATION>_<SEGMENT>/<ACTIVITY>`), one table at a time -- the
dominant per-segment cost of the reference.  `NpyResults` is a backend for the same protocols
(SupportsWriteTS, SupportsReadTS, SupportsWriteLogging) that needs no HDF5:

  * `write_ts(data_frame, category, operation, segment, activity)` stores one frame, like the HDF5 backend, so it can sit
    below the reference's own IOManager (`IOManager(NpyResults(dir))` as output backend);
  * `open_block(operation, activity, segments, columns, tindex)` creates ONE memory-mapped float32 array
    [interval][column][segment] for a whole batch -- the orientation the GPU delivers -- into which
    `stream_batch()` below copies each time chunk as soon as its device-to-host copy has finished, while the GPU
    already works on the next chunk (north_star subsystem 4: "output streaming of only the requested SAVE timeseries back
    to the IOManager in time chunks on a side stream");
  * `read_ts(category, operation, segment, activity)` gives any stored frame back as the DataFrame the HDF5 backend would
    return, whichever way it was written.

Layout on disk: `<root>/index.json` + one `.npy` per block / frame (numpy format, readable without this package).
"""
from __future__ import annotations

import json
import os

import numpy as np


def _cat(category):
    v = getattr(category, "value", category)
    return "INPUT" if str(v).upper().startswith("INPUT") else "RESULT"


class NpyResults:
    def __init__(self, root, mode="a"):
        self.root = str(root)
        os.makedirs(self.root, exist_ok=True)
        self._index_path = os.path.join(self.root, "index.json")
        self.index = {"frames": {}, "blocks": {}}
        if mode != "w" and os.path.exists(self._index_path):
            with open(self._index_path) as f:
                self.index = json.load(f)
        self._open = {}
        self.log = None
        self.versions = None

    # ---- index ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _key(category, operation, segment, activity):
        return "%s/%s_%s/%s" % (_cat(category), operation, segment, activity)

    def flush(self):
        for a in self._open.values():
            a.flush()
        tmp = self._index_path + ".tmp"
        with open(tmp, "w") as f:
            json.dump(self.index, f)
        os.replace(tmp, self._index_path)

    def close(self):
        self.flush()
        self._open.clear()

    @staticmethod
    def _tindex_meta(tindex):
        import pandas as pd

        t = pd.DatetimeIndex(tindex)
        step = (t[1] - t[0]) if len(t) > 1 else pd.Timedelta(0)
        regular = len(t) < 3 or bool((np.diff(t.asi8) == np.diff(t.asi8)[0]).all())
        if regular:
            return {"start": str(t[0]), "step_ns": int(step.value), "n": len(t)}
        return {"stamps_ns": [int(x) for x in t.as_unit("ns").asi8]}

    @staticmethod
    def _tindex_from(meta):
        import pandas as pd

        if "stamps_ns" in meta:
            return pd.DatetimeIndex(np.array(meta["stamps_ns"], dtype="datetime64[ns]"))
        return pd.DatetimeIndex(pd.Timestamp(meta["start"]) + pd.to_timedelta(np.arange(meta["n"]) * meta["step_ns"], unit="ns"))

    # ---- SupportsWriteTS (hsp2io/protocols.py:35-43): one frame ---------------------------------------------------------
    def write_ts(self, data_frame, category, operation=None, segment=None, activity=None, *args, **kwargs):
        key = self._key(category, operation, segment, activity)
        fn = key.replace("/", "__") + ".npy"
        a = np.ascontiguousarray(data_frame.to_numpy())
        np.save(os.path.join(self.root, fn), a)
        self.index["frames"][key] = {"file": fn, "columns": [str(c) for c in data_frame.columns],
                                     "tindex": self._tindex_meta(data_frame.index)}

    # ---- a whole batch: [interval][column][segment], written in time chunks -------------------------------------------
    def open_block(self, operation, activity, segments, columns, tindex, dtype=np.float32):
        """-> writable memory map [len(tindex)][len(columns)][len(segments)]; rows [t0:t1] may be assigned in any order."""
        name = "%s__%s__block%d.npy" % (operation, activity, len(self.index["blocks"]))
        shape = (len(tindex), len(columns), len(segments))
        a = np.lib.format.open_memmap(os.path.join(self.root, name), mode="w+", dtype=np.dtype(dtype), shape=shape)
        self.index["blocks"][name] = {"operation": operation, "activity": activity, "segments": [str(s) for s in segments],
                                      "columns": [str(c) for c in columns], "tindex": self._tindex_meta(tindex)}
        self._open[name] = a
        return a

    # ---- SupportsReadTS (hsp2io/protocols.py:26-33) ---------------------------------------------------------------------
    def read_ts(self, category, operation=None, segment=None, activity=None):
        import pandas as pd

        key = self._key(category, operation, segment, activity)
        fr = self.index["frames"].get(key)
        if fr is not None:
            return pd.DataFrame(np.load(os.path.join(self.root, fr["file"])), index=self._tindex_from(fr["tindex"]), columns=fr["columns"])
        if _cat(category) == "RESULT":
            for name, b in self.index["blocks"].items():
                if b["operation"] == operation and b["activity"] == activity and str(segment) in b["segments"]:
                    a = self._open.get(name)
                    if a is None:
                        a = np.load(os.path.join(self.root, name), mmap_mode="r")
                    i = b["segments"].index(str(segment))
                    return pd.DataFrame(np.ascontiguousarray(a[:, :, i]), index=self._tindex_from(b["tindex"]), columns=b["columns"])
        return pd.DataFrame()  # hdf.py:91-92: an absent table reads as an empty frame

    # ---- SupportsWriteLogging (hsp2io/protocols.py:45-51) ---------------------------------------------------------------
    def write_log(self, hsp2_log):
        self.log = hsp2_log
        hsp2_log.to_csv(os.path.join(self.root, "logfile.csv"))

    def write_versioning(self, versions):
        self.versions = versions
        versions.to_csv(os.path.join(self.root, "versions.csv"))


def stream_batch(sink, batch, siminfo, operation, segment_names, save_tables, forcing_chunks, chunk, saveall=False, pinned=True):
    """Run `batch` (engine.LandBatch, layout "plain") over the whole run in time chunks and stream the SAVEd series into `sink`
    (NpyResults) as each chunk arrives: what the reference does with save_timeseries after every (segment, activity)
    (utilities.py:292-333: columns = SAVE keys ∩ produced series -- all of them under saveall, else those whose flag is
    set --, float32, sorted by name), for all segments at once and overlapped with the computation.

    forcing_chunks : callable (t0, t1) -> forcing [t1-t0][nf][n] in `batch.forcing_rows` / device segment order (C-contiguous,
                     batch.forcing_dtype), or None in shared-met mode
    save_tables    : {activity: SAVE dict}, the same for every segment of the batch
    -> {activity: (columns, block)}"""
    from .engine import pinned_empty

    if batch.layout != "plain":
        raise ValueError("stream_batch needs LandBatch(layout='plain')")
    tindex = siminfo["tindex"]
    steps = len(tindex)
    if batch.order is not None:
        # the store keeps the batch's DEVICE order (divergence-aware sorting): a block names its segments, so nothing is
        # permuted on the way out and read_ts finds a segment by name
        segment_names = [segment_names[i] for i in batch.order]
    rows_of = {}
    for r, full in enumerate(batch.save_rows):
        act, name = full.split("_", 1)
        rows_of.setdefault(act, {})[name] = r
    plan = {}
    for act, savedict in save_tables.items():
        have = rows_of.get(act, {})
        cols = sorted(k for k in savedict if k in have and (saveall or savedict[k]))
        if not cols:
            continue
        plan[act] = (cols, np.array([have[k] for k in cols]), sink.open_block(operation, act, segment_names, cols, tindex))
    if batch.out_dtype != np.float32:
        batch.set_output_dtype(np.float32)  # the cast save_timeseries applies (utilities.py:313), done by the kernel's store
    alloc = (lambda shape: pinned_empty(shape, np.float32)) if pinned else (lambda shape: np.empty(shape, np.float32))
    nbuf = 3  # chunk c computes / copies out while c-1 is being written and c-2's write finishes
    bufs = [alloc((min(chunk, steps), batch.ns, batch.n)) for _ in range(nbuf)]
    from concurrent.futures import ThreadPoolExecutor

    def drain(t0, t1, k):
        o = bufs[k][: t1 - t0]
        for act, (cols, rows, block) in plan.items():
            lo, hi = int(rows.min()), int(rows.max())
            if hi - lo + 1 == len(rows) and np.array_equal(rows, np.arange(lo, hi + 1)):
                block[t0:t1] = o[:, lo:hi + 1, :]  # the activity's columns are consecutive rows: one strided block copy
            else:
                for j, r in enumerate(rows):  # one [chunk][n] slab per saved column (numpy releases the GIL while it copies)
                    block[t0:t1, j, :] = o[:, r, :]

    writes = [None] * nbuf
    chunks = [(t0, min(steps, t0 + chunk)) for t0 in range(0, steps, chunk)]
    with ThreadPoolExecutor(2) as pool:
        for c, (t0, t1) in enumerate(chunks):
            k = c % nbuf
            if writes[k] is not None:
                writes[k].result()  # the host buffer is free once its previous chunk is in the sink
            f = forcing_chunks(t0, t1) if forcing_chunks is not None else None
            if f is None:
                from . import _lib

                batch._kernel()
                _lib.check(batch.lib.hsp2g_run_host_ld(batch.h, int(t0), int(t1 - t0), None, batch.n, bufs[k].ctypes.data, batch.n))
            else:
                batch.run_chunk(t0, f, bufs[k][: t1 - t0])
            if c >= 1:
                batch.sync_oldest()  # chunk c-1 has arrived; chunk c keeps the GPU and the bus busy meanwhile
                p0, p1 = chunks[c - 1]
                writes[(c - 1) % nbuf] = pool.submit(drain, p0, p1, (c - 1) % nbuf)
        batch.sync()
        p0, p1 = chunks[-1]
        writes[(len(chunks) - 1) % nbuf] = pool.submit(drain, p0, p1, (len(chunks) - 1) % nbuf)
        for w in writes:
            if w is not None:
                w.result()
    sink.flush()
    return {act: (cols, block) for act, (cols, _rows, block) in plan.items()}
