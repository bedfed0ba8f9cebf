"""Output side of the batched engine: hand the SAVEd series of every segment to an HSP2 IOManager exactly as the
reference's `save_timeseries` does (src/hsp2/hsp2/utilities.py:292-333) -- SURVEY.md 8f rank 1.

Per (segment, activity) the reference builds `DataFrame(index=tindex)` with one column per key of `SAVE ∩ ts`, casts
it to float32, sorts the columns by name and calls
    timeseries.write_ts(data_frame=df, save_columns=[k for k, v in SAVE.items() if v or saveall] (or all columns when
                        saveall), category=Category.RESULTS, operation=..., segment=..., activity=..., compress=..., outstep=...)
and any object with that `write_ts` (hsp2io/protocols.py:35-43, hsp2io/io.py:58-95; daily / monthly / yearly aggregation
for outstep 3/4/5 happens inside IOManager.write_ts) is a valid sink.  Here the float32 cast has already been done on the
GPU (hsp2g_set_output_dtype(HSP2G_F32): the kernel's store rounds to nearest even like `astype(np.float32)`), so only half
the bytes crossed PCIe and this module merely slices columns.
"""
from __future__ import annotations

import numpy as np


def result_category():
    """hsp2io's Category.RESULTS when the reference package is importable, else its value (protocols.py:10-12)."""
    try:
        from hsp2.hsp2io.protocols import Category  # type: ignore

        return Category.RESULTS
    except Exception:
        return "RESULT"


def save_batch(io_manager, siminfo, operation, segment_names, save_tables, saved_rows, out, saveall=False,
               compress=True, outstep=2, category=None):
    """Write the results of one batch.

    io_manager    : anything with the reference's write_ts(data_frame=, save_columns=, category=, operation=, segment=,
                    activity=, compress=, outstep=)
    siminfo       : needs 'tindex' (pandas DatetimeIndex of the run's intervals, main.py:94)
    segment_names : names of the batch's segments in row order
    save_tables   : {activity: SAVE dict of that activity} (the reference's ui['SAVE']: {series name: 0/1}), one for the
                    batch or a list with one per segment
    saved_rows    : LandBatch.save -- '<ACTIVITY>_<NAME>' of each row of `out`
    out           : [steps][len(saved_rows)][n] float32 (or float64, cast here like the reference does)
    """
    import pandas as pd

    category = result_category() if category is None else category
    tindex = siminfo["tindex"]
    out = np.asarray(out)
    if out.shape[0] != len(tindex):
        raise ValueError("out has %d intervals, tindex %d" % (out.shape[0], len(tindex)))
    rows_of = {}
    for r, full in enumerate(saved_rows):
        act, name = full.split("_", 1)
        rows_of.setdefault(act, {})[name] = r
    for i, seg in enumerate(segment_names):
        tabs = save_tables[i] if isinstance(save_tables, (list, tuple)) else save_tables
        for act, savedict in tabs.items():
            have = rows_of.get(act, {})
            cols = sorted(k for k in savedict if k in have)  # SAVE keys ∩ produced series, columns sorted (utilities.py:311-313)
            if not cols:
                print("DataFrame Empty for %s|%s|%s" % (operation, act, seg))  # utilities.py:332
                continue
            data = np.stack([out[:, have[k], i] for k in cols], axis=1).astype(np.float32, copy=False)
            df = pd.DataFrame(data, index=tindex, columns=cols)
            save_columns = list(df.columns) if saveall else [k for k, v in savedict.items() if v or saveall]
            io_manager.write_ts(data_frame=df, save_columns=save_columns, category=category, operation=operation,
                                segment=seg, activity=act, compress=compress, outstep=outstep)


def save_qual_batch(io_manager, siminfo, operation, segment_names, save_tables, qual_batch, out, saveall=False,
                    compress=True, outstep=2, category=None):
    """Results of a PQUAL / IQUAL batch (engine.QualBatch: rows = (segment, constituent)) the way save_timeseries writes them:
    for these two activities the reference collects every ts name that CONTAINS '_<key>' for a key of the SAVE table
    (utilities.py:294-301) -- so one frame per segment holds '<P|I>QUAL<k>_<NAME>' columns of all its constituents, and the key
    SOQO also pulls in SOQOC -- then casts to float32 and sorts the columns.

    save_tables : the activity's SAVE dict ({series name: 0/1}), one for the batch or a list with one per segment
    out         : [steps][len(qual_batch.save)][rows]
    """
    import pandas as pd

    category = result_category() if category is None else category
    tindex = siminfo["tindex"]
    out = np.asarray(out)
    if out.shape[0] != len(tindex):
        raise ValueError("out has %d intervals, tindex %d" % (out.shape[0], len(tindex)))
    act = "PQUAL" if operation == "PERLND" else "IQUAL"
    names = [full.split("_", 1)[1] for full in qual_batch.save]
    rows_of = {}
    for j, (i, k) in enumerate(qual_batch.rows):
        rows_of.setdefault(i, []).append((k, j))
    for i, seg in enumerate(segment_names):
        savedict = save_tables[i] if isinstance(save_tables, (list, tuple)) else save_tables
        cols = {}
        for k, j in rows_of.get(i, ()):
            for r, nm in enumerate(names):
                z = "%s%d_%s" % (act, k, nm)
                if any(("_" + y) in z for y in savedict):
                    cols[z] = out[:, r, j]
        if not cols:
            print("DataFrame Empty for %s|%s|%s" % (operation, act, seg))
            continue
        order = sorted(cols)
        df = pd.DataFrame(np.stack([cols[z] for z in order], axis=1).astype(np.float32, copy=False), index=tindex, columns=order)
        save_columns = list(df.columns) if saveall else [k for k, v in savedict.items() if v or saveall]
        io_manager.write_ts(data_frame=df, save_columns=save_columns, category=category, operation=operation, segment=seg,
                            activity=act, compress=compress, outstep=outstep)


# ---------------------------------------------------------------------------------------------------------------------------
# outstep 3 / 4 / 5: the period statistics IOManager.write_ts computes with pandas (hsp2io/io.py:74-94), on the device
# ---------------------------------------------------------------------------------------------------------------------------
OUTSTEP_FREQ = {3: "D", 4: "ME", 5: "YE"}  # io.py:76,83,90 ('M' and 'Y' in the pandas the reference was written for)


def period_bins(tindex, outstep):
    """-> (bin_of_step int32[len(tindex)], labels): which period of the reference's `resample(freq, origin='start')` every
    interval falls in (hsp2io/io.py:76-92).

    outstep 3 (daily): in the pandas the reference pins (1.2 / 2.x; environment.yml) 'D' is a fixed 24-hour Tick, for which
      origin='start' anchors the bins at the FIRST label of the frame: period k holds the labels in
      [tindex[0] + k days, tindex[0] + (k+1) days) and is labelled tindex[0] + k days.  An hourly run from midnight has its
      first label at 01:00, so every day holds 24 values and is labelled 01:00 -- not the calendar days that pandas 3 (where
      'D' became a calendar offset and ignores origin=) would give.  Computed here directly, identical to
      `resample('24h', origin='start')` in every pandas (tests/test_results.py pins that).
    outstep 4 / 5 (month / year ends): calendar offsets, for which origin= never had an effect; asked of pandas itself."""
    import pandas as pd

    if outstep == 3:
        if len(tindex) == 0:
            return np.zeros(0, dtype=np.int32), tindex[:0]
        day = pd.Timedelta(days=1)
        bins = np.asarray((tindex - tindex[0]) // day, dtype=np.int32)
        return np.ascontiguousarray(bins), pd.DatetimeIndex(tindex[0] + day * np.arange(int(bins[-1]) + 1))
    sizes = pd.Series(np.zeros(len(tindex), dtype=np.int8), index=tindex).resample(OUTSTEP_FREQ[outstep]).size()
    return np.repeat(np.arange(len(sizes), dtype=np.int32), sizes.to_numpy()), sizes.index


def aggregate_host_series(out32, bin_of_step, nbins, device=0):
    """[steps][rows][n] float32 -> [nbins][rows][3 = last, sum, mean][n] float32 through hsp2g_aggregate_f32 (upload, one launch,
    download): the device-resident entry point is what a streaming run calls on its output slot before the copy-out, which then
    moves 3 / (intervals per period) of the bytes."""
    import ctypes as C

    from . import _lib
    from .engine import tile_series

    lib = _lib.load()
    out32 = np.asarray(out32)
    if out32.dtype != np.float32:
        raise ValueError("period statistics are defined on the float32 frames save_timeseries builds")
    steps, rows, n = out32.shape
    t = np.ascontiguousarray(tile_series(out32.astype(np.float64)).astype(np.float32))  # [tiles][steps][rows][32]
    ntiles = t.shape[0]
    res = np.empty((ntiles, nbins, rows, 3, 32), dtype=np.float32)
    d_src, d_dst = lib.hsp2g_device_alloc(device, t.nbytes), lib.hsp2g_device_alloc(device, res.nbytes)
    if not d_src or not d_dst:
        raise MemoryError(lib.hsp2g_last_error().decode())
    try:
        _lib.check(lib.hsp2g_memcpy_h2d(device, d_src, t.ctypes.data, t.nbytes))
        b = np.ascontiguousarray(bin_of_step, dtype=np.int32)
        _lib.check(lib.hsp2g_aggregate_f32(device, d_src, ntiles, steps, rows, b.ctypes.data_as(C.POINTER(C.c_int32)), int(nbins),
                                           d_dst, None))
        _lib.check(lib.hsp2g_memcpy_d2h(device, res.ctypes.data, d_dst, res.nbytes))  # synchronous: orders after the launch
    finally:
        lib.hsp2g_device_free(device, C.c_void_p(d_src))
        lib.hsp2g_device_free(device, C.c_void_p(d_dst))
    return np.ascontiguousarray(res.transpose(1, 2, 3, 0, 4).reshape(nbins, rows, 3, ntiles * 32)[:, :, :, :n])


def save_batch_aggregated(backend, labels, operation, segment_names, save_tables, saved_rows, agg, saveall=False, category=None):
    """Hand an output BACKEND (hsp2io's SupportsWriteTS below IOManager: write_ts(data_frame, category, operation, segment,
    activity), io.py:95) the frames IOManager.write_ts builds for outstep 3 / 4 / 5: the saved columns with the suffixes
    _last, _sum, _aver, in that block order (io.py:79-81).  agg: [nbins][len(saved_rows)][3][n] from aggregate_host_series."""
    import pandas as pd

    category = result_category() if category is None else category
    rows_of = {}
    for r, full in enumerate(saved_rows):
        act, name = full.split("_", 1)
        rows_of.setdefault(act, {})[name] = r
    for i, seg in enumerate(segment_names):
        tabs = save_tables[i] if isinstance(save_tables, (list, tuple)) else save_tables
        for act, savedict in tabs.items():
            have = rows_of.get(act, {})
            cols = sorted(k for k in savedict if k in have and (saveall or savedict[k]))  # io.py:70-72 drops the rest first
            if not cols:
                continue
            blocks, names = [], []
            for stat, suffix in ((0, "_last"), (1, "_sum"), (2, "_aver")):
                blocks += [agg[:, have[k], stat, i] for k in cols]
                names += [k + suffix for k in cols]
            backend.write_ts(pd.DataFrame(np.stack(blocks, axis=1), index=labels, columns=names), category, operation, seg, act)
