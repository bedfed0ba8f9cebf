"""Shared time base of a run: calendar flags and monthly-table interpolation, restated in numpy.

Host-side restatement of the reference's pandas glue (src/hsp2/hsp2/utilities.py:136-211 `hoursval`,
`hourflag`, `monthval`, `dayval`, `initm`; src/hsp2/hsp2/SNOW.py:600-609 `hour6flag`), with its exact
conventions:

  * every series is keyed by INTERVAL START: element i <-> time  start + i*delt ; the reference builds them
    on whole calendar years and truncates to [start, stop] inclusive, i.e. steps+1 elements;
  * `dofirst` forces element 0 of the CALENDAR-YEAR series (Jan-1 00:00 of start.year), not of the run;
  * monthly tables are pinned at 00:00 of the 1st of each month, interpolated linearly in time ONCE PER DAY
    (constant within the day), December running toward next January, and held constant after Dec-1 of
    stop.year (where the reference's series ends);
  * the interpolation is numpy's `slope*(x-x0)+y0`, slope=(y1-y0)/(x1-x0), with x the integer timestamp of
    the pandas index as float64.  pandas>=3 indexes in microseconds, pandas<3 (what the reference supports)
    in nanoseconds; the two differ by 1 ulp on ~7% of days, so the unit is a parameter (`time_unit`).

The device never sees per-step per-segment parameter series: it gets the 12 monthly values per segment and
the shared per-step pair (x-x0, x1-x0) + month index from `Calendar.day_interp()`, and evaluates the same
expression (compiled without FMA contraction), which is bit-identical to `dayval()` below.
"""
from __future__ import annotations

import numpy as np

# utilities.py:46-48  dry lapse rate by hour of day
LAPSE = np.array([0.0035, 0.0035, 0.0035, 0.0035, 0.0035, 0.0035, 0.0037, 0.0040, 0.0041, 0.0043, 0.0046,
                  0.0047, 0.0048, 0.0049, 0.0050, 0.0050, 0.0048, 0.0046, 0.0044, 0.0042, 0.0040, 0.0038,
                  0.0037, 0.0036])
# utilities.py:50  1 = "summer" albedo curve (Apr..Sep)
SEASONS = np.array([0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0], dtype=np.float64)

_UNIT_PER_DAY = {"ns": 86400e9, "us": 86400e6, "ms": 86400e3, "s": 86400.0}


def default_time_unit():
    """Unit of the DatetimeIndex the installed pandas would build in utilities.dayval (see module doc)."""
    try:
        import pandas as pd

        return str(pd.date_range("2000-01-01", periods=2, freq="MS").unit)
    except Exception:  # pandas absent or too old to have .unit -> the reference's supported behaviour
        return "ns"


def _to_dt64(x):
    if isinstance(x, np.datetime64):
        return x.astype("datetime64[m]")
    if hasattr(x, "to_datetime64"):  # pandas.Timestamp
        return x.to_datetime64().astype("datetime64[m]")
    return np.datetime64(str(x)).astype("datetime64[m]")


class Calendar:
    """Time base of one run (start, stop, delt minutes): `steps` intervals, series have steps+1 entries."""

    def __init__(self, start, stop, delt, time_unit=None):
        self.start = _to_dt64(start)
        self.stop = _to_dt64(stop)
        self.delt = int(delt)
        if self.delt <= 0 or (1440 % self.delt != 0):
            raise ValueError("delt must divide 1440 minutes (HSPF INDELT); got %r" % (delt,))
        span = int((self.stop - self.start) / np.timedelta64(1, "m"))
        if span <= 0 or span % self.delt:
            raise ValueError("stop - start must be a positive multiple of delt")
        self.steps = span // self.delt  # main.py:93-95: len(date_range(start, stop, freq=delt)[1:])
        self.time_unit = time_unit or default_time_unit()
        if self.time_unit not in _UNIT_PER_DAY:
            raise ValueError("time_unit must be one of %s" % sorted(_UNIT_PER_DAY))
        n = self.steps + 1
        t = self.start + np.arange(n) * np.timedelta64(self.delt, "m")
        self.times = t  # interval-start stamps
        day = t.astype("datetime64[D]")
        self._day = day
        self._minute_of_day = ((t - day) / np.timedelta64(1, "m")).astype(np.int64)
        mon = t.astype("datetime64[M]")
        self._month = (mon.astype(np.int64) % 12).astype(np.int64)  # 0..11
        self._year = t.astype("datetime64[Y]").astype(np.int64) + 1970
        self._year0 = int(self.start.astype("datetime64[Y]").astype(np.int64)) + 1970
        self._stop_year = int(self.stop.astype("datetime64[Y]").astype(np.int64)) + 1970
        self._jan1 = np.datetime64("%04d-01-01T00:00" % self._year0).astype("datetime64[m]")

    # ---- utilities.py:136-158 ------------------------------------------------------------------
    def hoursval(self, hours24, dofirst=False, lapselike=False):
        hours24 = np.asarray(hours24, dtype=np.float64)
        n = self.steps + 1
        mod = self._minute_of_day
        if self.delt <= 60:
            hour = mod // 60
            if lapselike:
                out = hours24[hour].copy()  # ffill within the hour
                if dofirst:
                    # hours[0]=1 is then forward-filled over the first hour of the calendar year
                    first = (self.times >= self._jan1) & (self.times < self._jan1 + np.timedelta64(60, "m"))
                    out[first] = 1.0
            else:
                on_hour = (mod % 60) == 0
                out = np.where(on_hour, hours24[hour], 0.0)
                if dofirst:
                    out[self.times == self._jan1] = 1.0
        else:
            k = self.delt // 60  # whole hours per interval; bins are anchored at midnight
            h0 = mod // 60
            idx = (h0[:, None] + np.arange(k)[None, :]) % 24
            vals = hours24[idx]
            if dofirst:
                firstbin = (self.times <= self._jan1) & (self._jan1 < self.times + np.timedelta64(self.delt, "m"))
                vals = vals.copy()
                vals[firstbin, 0] = 1.0
            if lapselike:  # pandas' group mean: Kahan-compensated running sum / count
                acc = np.zeros(n)
                comp = np.zeros(n)
                for j in range(k):
                    y = vals[:, j] - comp
                    tt = acc + y
                    comp = (tt - acc) - y
                    acc = tt
                out = acc / k
            else:
                out = vals.max(axis=1)
        assert out.shape[0] == n
        return out.astype(np.float64)

    def hourflag(self, hourfg, dofirst=False):
        h = np.zeros(24)
        h[hourfg] = 1.0
        return self.hoursval(h, dofirst)

    def hour6flag(self, dofirst=False):  # SNOW.py:600-609
        h = np.zeros(24)
        h[0:6] = 1.0
        return self.hoursval(h, dofirst)

    def lapse(self):  # ATEMP.py:19
        return self.hoursval(LAPSE, lapselike=True)

    # ---- utilities.py:168-183 ------------------------------------------------------------------
    def monthval(self, monthly):
        monthly = np.asarray(monthly, dtype=np.float64)
        return monthly[self._month].astype(np.float64)

    # ---- utilities.py:186-202 ------------------------------------------------------------------
    def day_interp(self):
        """Per-step shared interpolation record: (m0, m1, xm, dxm, newday).

        value(step) = tab[m0] if xm == 0 else ((tab[m1]-tab[m0]) / dxm) * xm + tab[m0]
        xm  = (day - first_of_month) in `time_unit`; dxm = month length in `time_unit`.
        After Dec-1 of stop.year the reference's series has ended: value = tab[11] (m0=m1=11, xm=0).
        newday[i] is 1 where step i starts a new day relative to step i-1 (and at i=0).
        """
        upd = _UNIT_PER_DAY[self.time_unit]
        day = self._day
        mon0 = day.astype("datetime64[M]")
        first = mon0.astype("datetime64[D]")
        nxt = (mon0 + np.timedelta64(1, "M")).astype("datetime64[D]")
        xm = (day - first).astype(np.int64).astype(np.float64) * upd
        dxm = (nxt - first).astype(np.int64).astype(np.float64) * upd
        m0 = self._month.copy()
        m1 = (m0 + 1) % 12
        tail = (self._year == self._stop_year) & (m0 == 11)
        m1[tail] = 11
        xm[tail] = 0.0
        newday = np.ones(self.steps + 1, dtype=np.int64)
        newday[1:] = (day[1:] != day[:-1]).astype(np.int64)
        return m0, m1, xm, dxm, newday

    def dayval(self, monthly):
        monthly = np.asarray(monthly, dtype=np.float64)
        m0, m1, xm, dxm, _ = self.day_interp()
        y0, y1 = monthly[m0], monthly[m1]
        with np.errstate(invalid="ignore", divide="ignore"):
            slope = (y1 - y0) / dxm
            v = slope * xm + y0
        return np.where(xm == 0.0, y0, v)

    # ---- utilities.py:205-211 ------------------------------------------------------------------
    def initm(self, uci, flag, monthly_key, default):
        if flag and monthly_key in uci:
            return self.dayval(list(uci[monthly_key].values()))
        return np.full(self.steps, float(default))

    def initmdiv(self, uci, flag, monthly1, monthly2, default1, default2):
        """utilities.py:214-225: a parameter that is the month-by-month ratio of two tables (ACQOP / SQOLIM), or of the two
        constants.  Like the reference, only the SECOND table's presence is tested."""
        if flag and monthly1 and monthly2 in uci:
            m1, m2 = list(uci[monthly1].values()), list(uci[monthly2].values())
            return self.dayval([m1[m] / m2[m] for m in range(12)])
        return np.full(self.steps, default1 / default2)

    @classmethod
    def from_siminfo(cls, siminfo, time_unit=None):
        return cls(siminfo["start"], siminfo["stop"], siminfo["delt"], time_unit)
