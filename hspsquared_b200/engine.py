"""LandBatch: one batch of land segments resident on one B200, advanced in time chunks by the fused CUDA kernels.

This is the batched engine API: it bypasses the reference's per-segment Python loop
(main.py:91-275) but runs the same per-interval arithmetic for every segment (libhsp2g, include/hsp2g.h).
PyTorch, when present, is used by callers for pinned host memory / device tensors / streams only; this module needs
numpy and ctypes alone.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .calendar import Calendar
from .segstore import SegmentStore, forcing_fill

_dp = C.POINTER(C.c_double)

SECTION_OF = {"PERLND": ("ATEMP", "SNOW", "PWATER", "SEDMNT", "PSTEMP", "PWTGAS"), "IMPLND": ("ATEMP", "SNOW", "IWATER")}
ERRORS_OF = {"SNOW": ("SNOW0",), "PWATER": tuple("PWATER%d" % i for i in range(10)),
             "PSTEMP": tuple("PSTEMP%d" % i for i in range(6)), "IWATER": ("IWATER0",), "ATEMP": (), "SEDMNT": (),
             "PWTGAS": ()}


def _ptr(a, typ=_dp):
    return a.ctypes.data_as(typ)


def output_names(operation, act):
    """All '<SECTION>_<NAME>' output ids the active sections of this operation produce, in id order."""
    secs = [s for s in SECTION_OF[operation] if act & _lib.ACT_BITS[s]]
    return [n for n in _lib.names("outputs") if n.split("_", 1)[0] in secs]


class LandBatch:
    def __init__(self, store: SegmentStore, calendar: Calendar, forcing_names, save, device=0, timing=False):
        """store: SegmentStore; forcing_names: ordered forcing ids the forcing arrays carry (row order);
        save: ordered output ids '<SECTION>_<NAME>' to write (row order), or 'all'."""
        lib = _lib.load()
        self.lib = lib
        self.store = store
        self.cal = calendar
        self.n = store.n
        self.device = device
        self.h = _lib._h()
        op = _lib.PERLND if store.operation == "PERLND" else _lib.IMPLND
        _lib.check(lib.hsp2g_batch_create(device, op, self.n, store.act, float(calendar.delt), C.byref(self.h)))
        self.npad = int(lib.hsp2g_batch_npad(self.h))
        try:
            self._upload(forcing_names, save, timing)
        except Exception:
            self.close()
            raise

    # ------------------------------------------------------------------------------------------------
    def _upload(self, forcing_names, save, timing):
        lib, st, cal = self.lib, self.store, self.cal
        _lib.check(lib.hsp2g_set_params(self.h, _ptr(st.par), st.n))
        _lib.check(lib.hsp2g_set_flags(self.h, _ptr(st.flags, C.POINTER(C.c_uint64))))
        for mid, tab in st.monthly.items():
            _lib.check(lib.hsp2g_set_monthly(self.h, int(mid), _ptr(tab), st.n))
        _lib.check(lib.hsp2g_set_state(self.h, _ptr(st.state), st.n))
        # calendar records
        m0, m1, xm, dxm, newday = cal.day_interp()
        fl = (cal.hoursval(np.ones(24), dofirst=True) != 0).astype(np.uint8) * _lib.CAL_HRFG
        fl |= (cal.hourflag(0, dofirst=True) != 0).astype(np.uint8) * _lib.CAL_DAYFG
        fl |= (cal.hourflag(1, dofirst=True) != 0).astype(np.uint8) * _lib.CAL_HR1FG
        fl |= (cal.hour6flag(dofirst=True) != 0).astype(np.uint8) * _lib.CAL_HR6IND
        from .calendar import SEASONS

        fl |= (cal.monthval(SEASONS) != 0).astype(np.uint8) * _lib.CAL_SEASON
        fl |= (newday != 0).astype(np.uint8) * _lib.CAL_NEWDAY
        lapse = np.ascontiguousarray(cal.lapse())
        u8 = C.POINTER(C.c_uint8)
        m0b, m1b = m0.astype(np.uint8), m1.astype(np.uint8)
        xm, dxm = np.ascontiguousarray(xm), np.ascontiguousarray(dxm)
        _lib.check(lib.hsp2g_set_calendar(self.h, len(fl), _ptr(fl, u8), _ptr(m0b, u8), _ptr(m1b, u8), _ptr(xm),
                                          _ptr(dxm), _ptr(lapse)))
        # forcing layout + fills
        F = _lib.index("forcings")
        self.forcing_names = tuple(forcing_names)
        for nme in self.forcing_names:
            if nme not in F:
                raise KeyError("unknown forcing series %r" % nme)
        fills, opts = forcing_fill(st.operation, st.act, self.forcing_names)
        fill = np.zeros(len(F))
        for nme, v in fills.items():
            if nme in F:
                fill[F[nme]] = v
        ids = np.array([F[nme] for nme in self.forcing_names], dtype=np.int32)
        _lib.check(lib.hsp2g_set_forcing_layout(self.h, len(ids), _ptr(ids, C.POINTER(C.c_int32)), _ptr(fill), opts))
        self.nf = len(ids)
        # SAVE set
        O = _lib.index("outputs")
        avail = output_names(st.operation, st.act)
        if save == "all":
            save = avail
        self.save = tuple(save)
        for nme in self.save:
            if nme not in avail:
                raise KeyError("output %r is not produced by the active sections" % nme)
        sid = np.array([O[nme] for nme in self.save], dtype=np.int32)
        _lib.check(lib.hsp2g_set_save(self.h, len(sid), _ptr(sid, C.POINTER(C.c_int32))))
        self.ns = len(sid)
        _lib.check(lib.hsp2g_set_timing(self.h, int(bool(timing))))

    # ------------------------------------------------------------------------------------------------
    def run_host(self, t0, forcing, out=None):
        """forcing: float64 [nsteps][nf][n] (host, C-contiguous).  Enqueues copy-in, kernel and copy-out; call
        sync() before reading `out` ([nsteps][ns][n], allocated here when None)."""
        forcing = np.ascontiguousarray(forcing, dtype=np.float64)
        nsteps = forcing.shape[0]
        if forcing.shape != (nsteps, self.nf, self.n):
            raise ValueError("forcing must be [nsteps][%d][%d]" % (self.nf, self.n))
        if out is None:
            out = np.empty((nsteps, self.ns, self.n), dtype=np.float64)
        elif out.shape != (nsteps, self.ns, self.n) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be C-contiguous float64 [nsteps][%d][%d]" % (self.ns, self.n))
        self._keep = (forcing, out)
        _lib.check(self.lib.hsp2g_run_host(self.h, int(t0), int(nsteps), forcing.ctypes.data, self.n, out.ctypes.data,
                                           self.n))
        return out

    def run_host_ptr(self, t0, nsteps, forcing_ptr, out_ptr, ld=None):
        """Raw-pointer variant for pinned buffers owned by the caller (e.g. torch pinned tensors)."""
        ld = self.n if ld is None else ld
        _lib.check(self.lib.hsp2g_run_host(self.h, int(t0), int(nsteps), forcing_ptr, ld, out_ptr, ld))

    def run_device(self, t0, nsteps, d_forcing, d_out, stream=0):
        """Device-resident pass: d_forcing [nsteps][nf][npad], d_out [nsteps][ns][npad] raw device pointers."""
        _lib.check(self.lib.hsp2g_run_device(self.h, int(t0), int(nsteps), d_forcing, d_out, stream or None))

    def sync(self):
        _lib.check(self.lib.hsp2g_sync(self.h))

    def run(self, forcing, chunk=None):
        """Whole series through host buffers in chunks of `chunk` intervals -> out [steps][ns][n]."""
        forcing = np.ascontiguousarray(forcing, dtype=np.float64)
        steps = forcing.shape[0]
        chunk = chunk or steps
        out = np.empty((steps, self.ns, self.n), dtype=np.float64)
        keep = []
        for t0 in range(0, steps, chunk):
            t1 = min(steps, t0 + chunk)
            f = forcing[t0:t1]
            o = out[t0:t1]
            keep.append((f, o))
            _lib.check(self.lib.hsp2g_run_host(self.h, t0, t1 - t0, f.ctypes.data, self.n, o.ctypes.data, self.n))
        self.sync()
        return out

    # ------------------------------------------------------------------------------------------------
    def errors(self):
        """{error id: int64[n]} for the counters of the active sections."""
        E = _lib.index("errors")
        buf = np.zeros((len(E), self.n), dtype=np.int64)
        _lib.check(self.lib.hsp2g_get_errors(self.h, _ptr(buf, C.POINTER(C.c_int64)), self.n))
        return {nme: buf[i] for nme, i in E.items()}

    def errors_for(self, section, seg=0):
        e = self.errors()
        return np.array([e[k][seg] for k in ERRORS_OF[section]], dtype=np.int64)

    def get_state(self):
        S = _lib.index("states")
        buf = np.zeros((len(S), self.n), dtype=np.float64)
        _lib.check(self.lib.hsp2g_get_state(self.h, _ptr(buf), self.n))
        return buf

    def set_state(self, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        _lib.check(self.lib.hsp2g_set_state(self.h, _ptr(state), self.n))

    def kernel_stats(self):
        n = C.c_int64(0)
        ms = C.c_double(0.0)
        _lib.check(self.lib.hsp2g_kernel_stats(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def kernel_name(self):
        return self.lib.hsp2g_kernel_name(self.h).decode()

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.hsp2g_batch_destroy(self.h)
            self.h = _lib._h()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def pack_forcings(series, names, n):
    """{name: float64[steps] or [steps][n]} -> [steps][len(names)][n] (shared series are broadcast)."""
    steps = len(next(iter(series.values())))
    out = np.empty((steps, len(names), n), dtype=np.float64)
    for r, nme in enumerate(names):
        a = np.asarray(series[nme], dtype=np.float64)
        out[:, r, :] = a[:steps, None] if a.ndim == 1 else a[:steps]
    return out
