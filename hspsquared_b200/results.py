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
