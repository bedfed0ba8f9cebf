"""WDM (Watershed Data Management) time-series decoder -- the data format on the input side of the land-segment path.

The reference decodes WDM files with per-value njit loops and stores every dataset in an HDF5 file that `get_timeseries`
later reads back segment by segment (hsp2tools/readWDM.py:29-164, helpers :166-319).  Here a file is decoded once, block by
block with numpy, straight into arrays that `ingest.MetStore` lays out as the shared-met store `[t][row][station]` of
`hsp2g_set_shared_forcing` -- no intermediate file.  Written from the WDM layout (512-word records of 4-byte words; a
dataset label record holds pointers to its search attributes and to its data groups; a data group starts with a packed date
word and is a chain of blocks, each a control word `nval | tstep | tcode | compression | quality` followed by `nval` values or,
compressed, by one value standing for `nval`), with the same conventions as the reference's decoder where the format leaves a
choice (blocks never span records; a group ends one TGROUP unit after its start; hour 24 rolls into the next day;
`tests/test_wdm.py` holds the two decoders against each other on every dataset of the reference's WDM files).
"""
import numpy as np

RECORD = 512
# attribute id -> (name, kind): readWDM.py:16-25 lists the ones the reference reads; kind I = int32 word, R = float32 word,
# S<n> = n characters packed four to a word
ATTRIBUTES = {1: ("TSTYPE", "S4"), 2: ("STAID", "S16"), 7: ("ELEV", "R"), 8: ("LATDEG", "R"), 9: ("LNGDEG", "R"),
              10: ("DESCRP", "S80"), 11: ("DAREA", "R"), 17: ("TCODE", "I"), 22: ("DCODE", "I"), 27: ("TSBYR", "I"),
              28: ("TSBMO", "I"), 29: ("TSBDY", "I"), 30: ("TSBHR", "I"), 32: ("TFILL", "R"), 33: ("TSSTEP", "I"),
              34: ("TGROUP", "I"), 45: ("STNAM", "S48"), 83: ("COMPFG", "I"), 84: ("TSFORM", "I"), 85: ("VBTIME", "I"),
              288: ("SCENARIO", "S8"), 289: ("CONSTITUENT", "S8"), 290: ("LOCATION", "S8"), 443: ("DATCRE", "S12"),
              444: ("DATMOD", "S12")}
ATTRIBUTE_DEFAULTS = {"TSBDY": 1, "TSBHR": 1, "TSBMO": 1, "TSBYR": 1900, "TFILL": -999.0}  # readWDM.py:89
# time-unit codes (TCODE, TGROUP and the block control word): 1 s, 2 min, 3 h, 4 d, 5 month, 6 year, 7 century
_FIXED = {1: np.timedelta64(1, "s"), 2: np.timedelta64(60, "s"), 3: np.timedelta64(3600, "s"), 4: np.timedelta64(86400, "s")}
_MONTHS = {5: 1, 6: 12, 7: 1200}


class WdmSeries:
    """One time-series dataset: `times` datetime64[s] (interval starts), `values` float64 (the file's float32 values
    widened exactly), `tcode` / `tstep` (the dataset's nominal step) and the decoded attributes."""

    def __init__(self, dsn, attrs, times, values):
        self.dsn, self.attrs, self.times, self.values = dsn, attrs, times, values
        self.tcode, self.tstep = int(attrs.get("TCODE", 0)), int(attrs.get("TSSTEP", 0))

    @property
    def step(self):
        """Nominal step as timedelta64[s], or None for month / year / century data."""
        return _FIXED[self.tcode] * self.tstep if self.tcode in _FIXED else None

    def __len__(self):
        return len(self.values)


def _add_units(t, tcode, k):
    """t (datetime64[s]) + k units of `tcode` (k scalar or integer array).  Month-based units keep the day-of-month offset
    and let it run over into the following month, as the reference's date arithmetic does (readWDM.py:205-233)."""
    if tcode in _FIXED:
        return t + k * _FIXED[tcode]
    m = t.astype("datetime64[M]")
    return (m + k * _MONTHS[tcode]).astype("datetime64[s]") + (t - m.astype("datetime64[s]"))


def _group_start(word):
    """Packed group date (year << 14 | month << 10 | day << 5 | hour) -> datetime64[s]; hour 24 is the next midnight."""
    y, mo, d, h = int(word) >> 14, (int(word) >> 10) & 0xF, (int(word) >> 5) & 0x1F, int(word) & 0x1F
    return (np.datetime64("%04d-%02d" % (y, mo), "M").astype("datetime64[s]")
            + np.timedelta64(d - 1, "D").astype("timedelta64[s]") + np.timedelta64(h, "h").astype("timedelta64[s]"))


class WdmFile:
    """Decoded directory of a WDM file; `series(dsn)` decodes one dataset, `dsns` lists the time-series datasets."""

    def __init__(self, path):
        self.path = path
        self.i = np.fromfile(path, dtype="<i4")
        self.f = self.i.view("<f4")
        if self.i.size < RECORD or self.i[0] != -998:
            raise ValueError("%s is not a WDM file (first word is not -998)" % path)
        nrec = int(self.i[28])
        labels = {}
        for rec in range(1, nrec):
            b = rec * RECORD
            if b + RECORD > self.i.size:
                break
            if self.i[b + 5] == 1 and self.i[b:b + 4].any():  # dataset type 1 = time series; free records are zeroed links
                labels[int(self.i[b + 4])] = b
        if len(labels) != int(self.i[31]):
            raise ValueError("%s: directory says %d time-series datasets, %d label records found"
                             % (path, int(self.i[31]), len(labels)))
        self._labels = labels
        self._cache = {}

    @property
    def dsns(self):
        return sorted(self._labels)

    def attributes(self, dsn):
        b = self._labels[dsn]
        i, f = self.i, self.f
        psa = int(i[b + 9])
        count = int(i[b + psa - 1]) if psa > 0 else 0
        out = dict(ATTRIBUTE_DEFAULTS)
        for k in range(count):
            aid, ptr = int(i[b + psa + 1 + 2 * k]), b + int(i[b + psa + 2 + 2 * k]) - 1
            if aid not in ATTRIBUTES:
                continue
            name, kind = ATTRIBUTES[aid]
            if kind == "I":
                out[name] = int(i[ptr])
            elif kind == "R":
                out[name] = float(f[ptr])
            else:
                nchar = int(kind[1:])
                out[name] = i[ptr:ptr + nchar // 4].tobytes().decode("latin-1").strip()
        return out

    def series(self, dsn):
        if dsn in self._cache:
            return self._cache[dsn]
        if dsn not in self._labels:
            raise KeyError("dataset %r not in %s (has %r)" % (dsn, self.path, self.dsns))
        b = self._labels[dsn]
        i, f = self.i, self.f
        attrs = self.attributes(dsn)
        pdat, pdatv = int(i[b + 10]), int(i[b + 11])
        pointers = i[b + pdat + 1:b + pdatv - 1]
        pointers = pointers[pointers != 0]
        tgroup = int(attrs["TGROUP"])
        t_parts, v_parts = [], []
        for p in pointers:
            rec, off = (int(p) >> 9) - 1, (int(p) & 0x1FF) - 1
            pos = rec * RECORD + off
            forward = int(i[rec * RECORD + 3])  # next record of this dataset's chain (1-based), 0 at the end
            now = _group_start(i[pos])
            end = _add_units(now, tgroup, 1)
            pos, off = pos + 1, off + 1
            while now < end:
                c = int(i[pos])
                nval, ltstep, ltcode, comp = (c >> 16) & 0xFFFF, (c >> 10) & 0x3F, (c >> 7) & 0x7, (c >> 5) & 0x3
                t_parts.append(_add_units(now, ltcode, np.arange(nval, dtype=np.int64) * ltstep))
                now = _add_units(now, ltcode, nval * ltstep)
                if comp == 1:
                    v_parts.append(np.full(nval, f[pos + 1], dtype=np.float32))
                    used = 2
                else:
                    v_parts.append(f[pos + 1:pos + 1 + nval])
                    used = 1 + nval
                pos, off = pos + used, off + used
                if off >= RECORD - 1:  # the chain continues after the four link words of the next record
                    rec, off = forward, 4
                    pos = (rec - 1) * RECORD + off
                    forward = int(i[(rec - 1) * RECORD + 3])
        if not t_parts:
            s = WdmSeries(dsn, attrs, np.empty(0, "datetime64[s]"), np.empty(0))
        else:
            s = WdmSeries(dsn, attrs, np.concatenate(t_parts).astype("datetime64[s]"),
                          np.concatenate(v_parts).astype(np.float64))
        self._cache[dsn] = s
        return s
