"""Forcing ingest: EXT SOURCES rows + decoded input series -> the shared-met store of `hsp2g_set_shared_forcing`.

The reference assembles forcings one segment at a time: for every EXT SOURCES row of the segment `get_timeseries` reads the
source series, multiplies it by MFACTOR, resamples it to the run step with pandas (`transform`) and sums rows that feed the
same target (utilities.py:274-290, :83-133).  Real watersheds share a handful of gages between thousands of segments, so
here every distinct (source, transformation) is resampled ONCE with numpy into `series[t][row][station]`, a segment keeps an
index into the stations plus its MFACTORs, and the kernels read `series[t][r][station[i]] * mfactor[r][i]` -- no per-segment
forcing array exists on the host, crosses PCIe or is read from HBM (`met=shared` kernels, include/hsp2g.h).

`resample()` restates `transform` for series with a regular step (what readWDM produces): the append-a-copy-of-the-last-point
rule (:93-95), pass-through at equal step (:97), SAME / LAST forward fill (:101, :124), DIV (:125), the default rule by
`flowtype` (:103-119), SUM / MEAN / MAX / MIN binning with pandas' compensated summation (:120-123), INTERPOLATE (:127) and the `[start:stop][0:steps]` trim (:132-133).  Bit-level agreement with the reference's
`transform` and `get_timeseries` is held in `tests/test_ingest.py`.
"""
from collections import namedtuple

import numpy as np

# utilities.py:23-36 -- member names treated as flows by the default transformation
FLOWTYPE = frozenset((
    "PREC WIND WINMOV SOLRAD PETINP POTEV SURLI IFWLI AGWLI SLSED IVOL ICON PRECIP SNOWF PRAIN SNOWE WYIELD MELT SUPY SURO "
    "IFWO AGWO PERO IGWI PET CEPE UZET LZET AGWET BASET TAET IFWI UZI INFIL PERC LZI AGWI SOHT IOHT AOHT POHT SODOXM SOCO2M "
    "IODOXM IOCO2M AODOXM AOCO2M PODOXM POCO2M").split())
# utilities.py:494-501: members whose subscript is dropped from the `ts` name
_NO_SUBSCRIPT = frozenset("GATMP PREC DTMPG WINMOV DSOLAR SOLRAD CLOUD PETINP IRRINP POTEV DEWTMP WIND IVOL IHEAT".split())

SourceRow = namedtuple("SourceRow", "svolno mfactor tran tmemn tmemsb", defaults=("",))


def target_name(tmemn, tmemsb=""):
    """`clean_name` (utilities.py:494-510) for the members the land segments read."""
    return tmemn if tmemn in _NO_SUBSCRIPT else "%s%s" % (tmemn, tmemsb)


def _bin_reduce(bins, values, nbins, how):
    """Per-bin SUM / MEAN / MAX / MIN of `values` whose (sorted) bin numbers are `bins`; empty bins give 0 for SUM and NaN
    otherwise.  SUM and MEAN follow pandas' group_sum / group_mean recurrence: compensated summation in arrival order."""
    inside = (bins >= 0) & (bins < nbins)
    bins, values = bins[inside], values[inside]
    count = np.bincount(bins, minlength=nbins)
    first = np.concatenate(([0], np.cumsum(count)[:-1]))
    has = count > 0
    if how in ("SUM", "MEAN"):
        s, c = np.zeros(nbins), np.zeros(nbins)
        for j in range(int(count.max()) if len(count) else 0):
            m = count > j
            y = values[first[m] + j] - c[m]
            t = s[m] + y
            c[m] = (t - s[m]) - y
            s[m] = t
        if how == "SUM":
            return s
        out = np.full(nbins, np.nan)
        out[has] = s[has] / count[has]
        return out
    out = np.full(nbins, np.nan)
    if has.any():
        red = np.maximum if how == "MAX" else np.minimum
        out[has] = red.reduceat(values, first[has])
    return out


def resample(times, values, step, name, how, calendar, tcode=None):
    """float64[steps]: the series (`times` datetime64, `values`; nominal `step` timedelta64, None for month / year data;
    `tcode` the WDM time-unit code of the step, 2 = minutes ... 6 = years) on the run's interval starts, as
    utilities.transform returns it for a series whose index carries that frequency."""
    times = np.asarray(times).astype("datetime64[s]")
    values = np.asarray(values, dtype=np.float64)
    start, stop = calendar.start.astype("datetime64[s]"), calendar.stop.astype("datetime64[s]")
    delt = np.timedelta64(int(calendar.delt) * 60, "s")
    steps = calendar.steps
    if len(times) == 0:
        raise ValueError("empty source series for %s" % name)
    month_based = {5: "M", 6: "Y", 7: "Y"}.get(tcode)
    if month_based is None and step is None:
        raise ValueError("a source without a regular step needs its time-unit code")
    if times[-1] < stop:  # :93-95
        times, values = np.append(times, stop), np.append(values, values[-1])
    grid = start + np.arange(steps) * delt
    how = (how or "").strip().upper()

    if month_based is None and step == delt:  # :97 -- no resampling, only the [start:stop] slice
        k = np.searchsorted(times, grid, side="left")
        if k[-1] >= len(times) or not np.array_equal(times[k], grid):
            raise ValueError("source series for %s does not cover the run on the run's own step" % name)
        return values[k]

    # pandas anchors the bins of a sub-daily resample at midnight of the first day; the run's interval starts must be bin labels
    anchor = times[0].astype("datetime64[D]").astype("datetime64[s]")
    if (start - anchor) % delt != np.timedelta64(0, "s"):
        raise ValueError("run start %s is not on the %d-minute grid anchored at %s" % (start, calendar.delt, anchor))
    g0 = anchor + ((times[0] - anchor) // delt) * delt  # label of the first bin

    def pick(v):  # resample(freq).ffill(): the last source point at or before each label (NaN before the first)
        k = np.searchsorted(times, grid, side="right") - 1
        out = np.where(k >= 0, v[np.maximum(k, 0)], np.nan)
        if grid[0] < g0:
            raise ValueError("source series for %s starts at %s, after the run start %s" % (name, times[0], start))
        return out

    coarser = month_based is not None or step > delt
    if how in ("SAME", "LAST"):
        return pick(values)
    if how == "DIV":
        if month_based is not None:
            raise ValueError("DIV of a monthly / yearly source is not defined by the reference (utilities.py:125)")
        return pick(values * (delt / step))
    if how == "":
        # :107-119.  The reference tests the letters 'M' and 'Y' in the frequency's repr, so a step counted in minutes
        # ('<15 * Minutes>') takes the monthly branch as well; kept, because results must match.
        month_like = month_based == "M" or tcode == 2
        if month_like or month_based == "Y" or coarser:
            if name in FLOWTYPE:
                ratio = 1.0 / 730.5 if month_like else 1.0 / 8766.0 if month_based == "Y" else delt / step
                return pick(ratio * values)
            return pick(values)
        how = "SUM" if name in FLOWTYPE else "MEAN"
    if grid[0] < g0:
        raise ValueError("source series for %s starts at %s, after the run start %s" % (name, times[0], start))
    b0 = int((start - g0) // delt)
    if how in ("SUM", "MEAN", "MAX", "MIN"):
        bins = ((times - g0) // delt).astype(np.int64) - b0
        return _bin_reduce(bins, values, steps, how)
    if how == "INTERPOLATE":
        if (step % delt if coarser else delt % step) != np.timedelta64(0, "s"):
            raise NotImplementedError("INTERPOLATE between steps that are not multiples of one another depends on the "
                                      "pandas version (whether off-grid source points take part); not restated")
        # upsample to the anchored grid (only source points that fall on labels survive), then linear interpolation by
        # POSITION, forward only: leading gaps stay NaN, trailing ones repeat the last value (pandas' defaults)
        on = (times - g0) % delt == np.timedelta64(0, "s")
        xp = ((times[on] - g0) // delt).astype(np.float64)
        x = ((grid - g0) // delt).astype(np.float64)
        out = np.interp(x, xp, values[on])
        out[x < xp[0]] = np.nan
        return out
    if how == "ZEROFILL":
        raise NotImplementedError("ZEROFILL: the reference's own expression (utilities.py:126) does not run under the pandas "
                                  "of this image, so there is nothing to pin it to")
    raise ValueError("unknown transformation %r for %s" % (how, name))


def _series_step(s):
    """(step, tcode) of a WdmSeries-like object: step None for month / year / century data."""
    tcode = getattr(s, "tcode", None)
    if tcode in (5, 6, 7):
        return None, tcode
    step = s.step
    return (np.timedelta64(step, "s") if not isinstance(step, np.timedelta64) else step.astype("timedelta64[s]")), tcode


class MetStore:
    """Shared-met layout of a set of segments: `series` [steps][rows][stations], `station` [n], `mfactor` [rows][n]."""

    def __init__(self, forcing_names, series, station, mfactor):
        self.forcing_names, self.series, self.station, self.mfactor = list(forcing_names), series, station, mfactor

    @property
    def n_stations(self):
        return self.series.shape[2]

    @classmethod
    def build(cls, calendar, forcing_names, segment_rows, read_series):
        """segment_rows: per segment, its EXT SOURCES rows (SourceRow or tuples (svolno, mfactor, tran, tmemn[, tmemsb]));
        read_series(svolno) -> object with .times, .values and .tcode/.tstep (wdm.WdmSeries) -- called once per source.

        A target fed by one row whose transformation only forward-fills raw values (SAME, LAST, equal step) keeps its
        MFACTOR as a per-segment factor: series * MFACTOR is then the very product the reference forms.  Every other case
        (DIV, aggregations, several rows summed into one target) is evaluated on the host in the reference's order
        -- (x * MFACTOR) * ratio, row after row -- and stored as its own station column with factor 1."""
        names = list(forcing_names)
        n = len(segment_rows)
        sources, columns = {}, {}

        def source(svolno):
            if svolno not in sources:
                sources[svolno] = read_series(svolno)
            return sources[svolno]

        def fill_only(row):
            step, tcode = _series_step(source(row.svolno))
            how = (row.tran or "").strip().upper()
            delt = np.timedelta64(int(calendar.delt) * 60, "s")
            if step is not None and step == delt:
                return True
            if how in ("SAME", "LAST"):
                return True
            return how == "" and row.tmemn not in FLOWTYPE and (step is None or tcode == 2 or step > delt)

        def column(name, key):
            if (name, key) not in columns:
                total = None
                for svolno, mfactor, tran, tmemn in key:
                    s = source(svolno)
                    step, tcode = _series_step(s)
                    v = s.values if mfactor == 1.0 else s.values * mfactor  # utilities.py:281-282
                    t = resample(s.times, v, step, tmemn, tran, calendar, tcode)
                    total = t if total is None else total + t  # :286-289
                columns[(name, key)] = total
            return columns[(name, key)]

        station_of, stations = {}, []
        station = np.empty(n, dtype=np.int32)
        mfac = np.ones((len(names), n))
        for i, rows in enumerate(segment_rows):
            rows = [r if isinstance(r, SourceRow) else SourceRow(*r) for r in rows]
            by_target = {}
            for r in rows:
                by_target.setdefault(target_name(r.tmemn, r.tmemsb), []).append(r)
            keys = []
            for j, nm in enumerate(names):
                rs = by_target.get(nm)
                if not rs:
                    raise ValueError("segment %d has no EXT SOURCES row for %s" % (i, nm))
                if len(rs) == 1 and fill_only(rs[0]):
                    keys.append(((rs[0].svolno, 1.0, rs[0].tran, rs[0].tmemn),))
                    mfac[j, i] = rs[0].mfactor
                else:
                    keys.append(tuple((r.svolno, float(r.mfactor), r.tran, r.tmemn) for r in rs))
            keys = tuple(keys)
            if keys not in station_of:
                station_of[keys] = len(stations)
                stations.append(keys)
            station[i] = station_of[keys]
        series = np.empty((calendar.steps, len(names), len(stations)))
        for k, keys in enumerate(stations):
            for j, nm in enumerate(names):
                series[:, j, k] = column(nm, keys[j])
        return cls(names, series, station, mfac)

    def expand(self, segments=None):
        """[steps][rows][n] per-segment forcing arrays (what get_timeseries would have built) -- for checks and for the
        per-segment-array kernels; the shared-met kernels never materialise this."""
        st = self.station if segments is None else self.station[segments]
        mf = self.mfactor if segments is None else self.mfactor[:, segments]
        return self.series[:, :, st] * mf[None, :, :]

    def apply(self, batch):
        """Hand the store to a LandBatch (hsp2g_set_shared_forcing); batch.run_shared() then needs no forcing arrays."""
        if list(batch.forcing_names) != self.forcing_names:
            raise ValueError("the batch was built for forcings %r, the store holds %r" % (batch.forcing_names, self.forcing_names))
        batch.set_shared_forcing(self.series, self.station, self.mfactor)
