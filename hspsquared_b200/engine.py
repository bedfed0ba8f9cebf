"""LandBatch: one batch of land segments resident on one B200, advanced in time chunks by the fused CUDA kernels.

This is the batched engine API: it bypasses the reference's per-segment Python loop
(main.py:91-275) but runs the same per-interval arithmetic for every segment (libhsp2g, include/hsp2g.h).
PyTorch, when present, is used by callers for pinned host memory / device tensors / streams only; this module needs
numpy and ctypes alone.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref

import numpy as np

from . import _lib
from .calendar import Calendar
from .segstore import SegmentStore, forcing_fill

_dp = C.POINTER(C.c_double)

SECTION_OF = {"PERLND": ("ATEMP", "SNOW", "PWATER", "SEDMNT", "PSTEMP", "PWTGAS", "PQUAL"),
              "IMPLND": ("ATEMP", "SNOW", "IWATER", "SOLIDS", "IWTGAS", "IQUAL")}
ERRORS_OF = {"SNOW": ("SNOW0",), "PWATER": tuple("PWATER%d" % i for i in range(10)),
             "PSTEMP": tuple("PSTEMP%d" % i for i in range(6)), "IWATER": ("IWATER0",), "ATEMP": (), "SEDMNT": (),
             "PWTGAS": (), "SOLIDS": (), "IWTGAS": (), "PQUAL": ("PQUAL0", "PQUAL1"), "IQUAL": ("IQUAL0", "IQUAL1")}


def _ptr(a, typ=_dp):
    return a.ctypes.data_as(typ)


def output_names(operation, act):
    """All '<SECTION>_<NAME>' output ids the active sections of this operation produce, in id order."""
    secs = [s for s in SECTION_OF[operation] if act & _lib.ACT_BITS[s]]
    return [n for n in _lib.names("outputs") if n.split("_", 1)[0] in secs]


def pinned_empty(shape, dtype=np.float64, write_combined=False):
    """numpy array over page-locked host memory (cudaHostAlloc through the C ABI): copies from / to it run asynchronously at
    PCIe speed.  write_combined=True for buffers the host only WRITES (forcings); reading such memory from the CPU is slow."""
    import weakref

    lib = _lib.load()
    dtype = np.dtype(dtype)
    nbytes = max(1, int(np.prod(shape)) * dtype.itemsize)
    ptr = lib.hsp2g_host_alloc_flags(nbytes, _lib.HOST_WRITE_COMBINED if write_combined else 0)
    if not ptr:
        raise MemoryError(lib.hsp2g_last_error().decode())
    raw = (C.c_char * nbytes).from_address(ptr)
    weakref.finalize(raw, lib.hsp2g_host_free, ptr)
    return np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class _PageLocked:
    """Page-lock ordinary numpy arrays in place (cudaHostRegister) so that their copies run asynchronously at PCIe speed, where
    the driver stages a pageable array through a bounce buffer at a tenth of it.  Measured on the B200 hosts
    (profiles/r2t_run_phases_ordinary_arrays.json): registering the 2.1 + 4.2 GB of one headline chunk costs 0.45-0.8 s, four
    to seven times the copies and the kernel it serves -- so a registration is KEPT for as long as the array lives (a caller who
    hands run() the same forcing array or the same out= array again pays it once) and undone by a finalizer when the array is
    collected, before its memory is returned.  Arrays that already are page-locked (pinned_empty, a caller's own registration)
    are left alone.  HSP2G_KEEP_REGISTERED=0: register for the duration of the call only."""

    MIN_BYTES = 1 << 20
    KEEP = os.environ.get("HSP2G_KEEP_REGISTERED", "1") != "0"
    _kept = {}  # (address, bytes) -> weakref.finalize of the array that owns the registration
    _lock = threading.Lock()

    def __init__(self, lib, arrays):
        self.lib, self.ptrs = lib, []
        for a in arrays:
            if a is None or a.nbytes < self.MIN_BYTES:
                continue
            key = (a.ctypes.data, a.nbytes)
            with self._lock:
                fin = self._kept.get(key)
                if fin is not None and fin.alive:
                    continue  # still registered from an earlier call
                if lib.hsp2g_host_register(a.ctypes.data, a.nbytes) != 0:
                    continue  # refused (already page-locked by somebody else, or not lockable): the copies are staged
                if self.KEEP:
                    fin = self._kept[key] = weakref.finalize(a, self._forget, lib, key)
                    fin.atexit = False  # at interpreter exit the driver unpins everything itself
                else:
                    self.ptrs.append(a.ctypes.data)

    @classmethod
    def _forget(cls, lib, key):
        with cls._lock:
            cls._kept.pop(key, None)
        lib.hsp2g_host_unregister(key[0])

    def release(self):
        for ptr in self.ptrs:
            self.lib.hsp2g_host_unregister(ptr)
        self.ptrs = []


def _prefault(a, threads=None):
    """Touch the pages of a freshly allocated array from several threads: the kernel zeroes a page in the thread that first
    touches it (1.5-5 GB/s from one thread, several times that from eight), and cudaHostRegister would do it alone."""
    v = a.reshape(-1).view(np.uint8)
    if v.size < (64 << 20):
        return
    from concurrent.futures import ThreadPoolExecutor

    threads = threads or min(8, os.cpu_count() or 1)
    cuts = np.linspace(0, v.size, threads * 4 + 1).astype(np.int64) // 4096 * 4096
    cuts[-1] = v.size
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda i: v[cuts[i]:cuts[i + 1]:4096].fill(0), range(len(cuts) - 1)))


class LandBatch:
    def __init__(self, store: SegmentStore, calendar: Calendar, forcing_names, save, device=0, timing=False, order="flags",
                 layout="plain", jit=None):
        """store: SegmentStore; forcing_names: ordered forcing ids the arrays given to run() carry (row order);
        save: ordered output ids '<SECTION>_<NAME>' run() returns (row order), or 'all'.  The raw series buffers of
        run_host / run_device use `forcing_rows` / `save_rows` (the same names in ascending id order).
        order: "flags" (default: SegmentStore.divergence_order(), the identity when every segment carries the same option
        word), None, or a permutation: device position i holds the caller's segment order[i].  run(), errors(), get_state()
        and set_state() keep speaking the caller's order; the raw buffers of run_host / run_device are in device order
        (`self.order`).
        layout: "plain" (default) -- series are [interval][row][segment], the orientation of the reference's ts arrays, on
        both sides of PCIe and in HBM (include/hsp2g.h HSP2G_PLAIN) -- or "tiled" ([tile][interval][row][32]).
        jit: build a kernel for exactly this batch (hspsquared_b200/jit.py); None = for batches of at least jit.MIN_SEGMENTS
        segments, unless HSP2G_JIT=0."""
        lib = _lib.load()
        self.lib = lib
        if isinstance(order, str) and order == "flags" and bool((store.flags == store.flags[0]).all()):
            order = None  # one option word: nothing to sort
        if order is not None:
            import copy as _copy

            store = _copy.copy(store)  # the caller's store keeps its order
            store.monthly = dict(store.monthly)
            perm = store.divergence_order() if isinstance(order, str) and order == "flags" else np.asarray(order, dtype=np.int64)
            store.permute(perm)
            self.order = perm
            self._inv = np.argsort(perm)
        else:
            self.order = None
            self._inv = None
        self.store = store
        self.cal = calendar
        self.n = store.n
        self.device = device
        self._keep = []
        self.out_dtype = np.dtype(np.float64)
        self.forcing_dtype = np.dtype(np.float64)
        self.shared = False
        self.met_table = False
        self.linked = False  # the last launch read its forcing rows in place from another batch's buffers (run_device_linked)
        if layout not in ("plain", "tiled"):
            raise ValueError("layout must be 'plain' or 'tiled'")
        self.layout = layout
        from . import jit as _jit

        self._jit = _jit
        self.jit = _jit.enabled(store.n) if jit is None else bool(jit)
        self.jit_word = True  # fix the batch-wide option word at compile time when there is one
        self.jit_error = None
        self.jit_meta = None
        self._kernel_for = None
        self.h = _lib._h()
        op = _lib.PERLND if store.operation == "PERLND" else _lib.IMPLND
        _lib.check(lib.hsp2g_batch_create(device, op, self.n, store.act, float(calendar.delt), C.byref(self.h)))
        self.npad = int(lib.hsp2g_batch_npad(self.h))
        try:
            self._upload(forcing_names, save, timing)
        except Exception:
            self.close()
            raise

    @staticmethod
    def climate_order(store, forcing_names, forcing_sample, bins=16):
        """Divergence-aware segment order from the forcings themselves: option word first (SegmentStore.divergence_order), then
        a climate key read off a sample of the batch's own forcing series ([steps][rows][n], rows = forcing_names; a few days
        are enough) -- the segment's mean air temperature in `bins` classes, then its mean precipitation.  A warp is 32
        consecutive segments: lanes that see similar weather build, melt and lose their snowpack in the same intervals and
        get wet together, so they take the same branches.  Pass the result as LandBatch(order=...); run() and the state /
        error accessors keep speaking the caller's order."""
        return store.divergence_order(keys=LandBatch.climate_keys(forcing_names, forcing_sample, bins))

    @staticmethod
    def climate_keys(forcing_names, forcing_sample, bins=16):
        """The per-segment sort keys climate_order uses (most significant first): air-temperature class, mean precipitation."""
        names = list(forcing_names)
        f = np.asarray(forcing_sample)
        keys = []
        tname = "AIRTMP" if "AIRTMP" in names else ("GATMP" if "GATMP" in names else None)
        if tname is not None:
            t = f[:, names.index(tname), :].mean(axis=0)
            span = float(t.max() - t.min())
            keys.append(np.floor((t - t.min()) / span * bins).clip(0, bins - 1) if span > 0 else np.zeros_like(t))
        if "PREC" in names:
            keys.append(f[:, names.index("PREC"), :].mean(axis=0))
        return tuple(keys)

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
        # The caller's row order is kept for run(); the DEVICE row order is ascending id, which is what the kernels
        # specialised at compile time for a row set expect (include/hsp2g.h hsp2g_set_forcing_layout).
        F = _lib.index("forcings")
        self.forcing_names = tuple(forcing_names)
        for nme in self.forcing_names:
            if nme not in F:
                raise KeyError("unknown forcing series %r" % nme)
        self._fperm = sorted(range(len(self.forcing_names)), key=lambda i: F[self.forcing_names[i]])
        self.forcing_rows = tuple(self.forcing_names[i] for i in self._fperm)
        fills, opts = forcing_fill(st.operation, st.act, self.forcing_names)
        fill = np.zeros(len(F))
        for nme, v in fills.items():
            if nme in F:
                fill[F[nme]] = v
        ids = np.array([F[nme] for nme in self.forcing_rows], dtype=np.int32)
        if getattr(st, "units", 1) == 2:
            opts |= _lib.OPT_METRIC  # the store's parameters and states are already in English units (segstore.build)
        self.opts = opts
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
        sperm = sorted(range(len(self.save)), key=lambda i: O[self.save[i]])
        self.save_rows = tuple(self.save[i] for i in sperm)
        self._sinv = np.argsort(np.array(sperm, dtype=np.int64)) if sperm else np.zeros(0, dtype=np.int64)
        sid = np.array([O[nme] for nme in self.save_rows], dtype=np.int32)
        _lib.check(lib.hsp2g_set_save(self.h, len(sid), _ptr(sid, C.POINTER(C.c_int32))))
        self.ns = len(sid)
        _lib.check(lib.hsp2g_set_timing(self.h, int(bool(timing))))
        _lib.check(lib.hsp2g_set_series_layout(self.h, _lib.PLAIN if self.layout == "plain" else _lib.TILED))

    # ------------------------------------------------------------------------------------------------
    @property
    def ntiles(self):
        return self.npad // _lib.TILE

    def _kernel(self, linked=False):
        """Before every launch: make sure the kernel built for this batch matches its current configuration (output and
        forcing element types, layout, shared-met mode and linked forcing rows can change after construction)."""
        self.linked = bool(linked)
        if not self.jit:
            return
        key = (self.out_dtype.itemsize, self.forcing_dtype.itemsize, self.layout, self.shared, self.met_table, self.jit_word,
               self.linked)
        if key != self._kernel_for:
            self._kernel_for = key
            self._jit.specialize(self, word=self.jit_word)

    def set_layout(self, layout):
        if layout not in ("plain", "tiled"):
            raise ValueError("layout must be 'plain' or 'tiled'")
        _lib.check(self.lib.hsp2g_set_series_layout(self.h, _lib.PLAIN if layout == "plain" else _lib.TILED))
        self.layout = layout

    def set_forcing_dtype(self, dtype):
        """np.float64 (default) or np.float32: element type of the forcing series given to run() / run_host / run_device.
        float32 is exact whenever the source is float32 (WDM, HDF5 /TIMESERIES; utilities.py:283-287) and halves the bytes."""
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("forcing dtype must be float64 or float32")
        _lib.check(self.lib.hsp2g_set_forcing_dtype(self.h, _lib.F32 if dtype == np.float32 else _lib.F64))
        self.forcing_dtype = dtype

    def run_host(self, t0, forcing_tiled, out_tiled=None):
        """forcing_tiled: float64 [ntiles][nsteps][nf][32] (host, C-contiguous, see tile_series) with rows in
        `self.forcing_rows` order; saved rows come back in `self.save_rows` order.  Enqueues copy-in,
        kernel and copy-out; call sync() before reading `out_tiled` ([ntiles][nsteps][ns][32], allocated when None)."""
        f = forcing_tiled
        if f.dtype != self.forcing_dtype or not f.flags.c_contiguous or f.ndim != 4 or f.shape[0] != self.ntiles or \
                f.shape[2:] != (self.nf, _lib.TILE):
            raise ValueError("forcing must be C-contiguous %s [%d][nsteps][%d][%d]" % (self.forcing_dtype, self.ntiles, self.nf, _lib.TILE))
        nsteps = f.shape[1]
        shape_o = (self.ntiles, nsteps, self.ns, _lib.TILE)
        if out_tiled is None:
            out_tiled = np.empty(shape_o, dtype=self.out_dtype)
        elif out_tiled.shape != shape_o or out_tiled.dtype != self.out_dtype or not out_tiled.flags.c_contiguous:
            raise ValueError("out must be C-contiguous %s %r" % (self.out_dtype, shape_o))
        if self.layout != "tiled":
            raise ValueError("run_host takes tile-major buffers: LandBatch(layout='tiled'); run_chunk is the plain-layout call")
        self._keep.append((f, out_tiled))
        self._kernel()
        _lib.check(self.lib.hsp2g_run_host(self.h, int(t0), int(nsteps), f.ctypes.data, out_tiled.ctypes.data))
        return out_tiled

    def run_host_ptr(self, t0, nsteps, forcing_ptr, out_ptr, ldf=None, ldo=None):
        """Raw-pointer variant for host buffers owned by the caller (pinned_empty, torch pinned tensors): tile-major buffers,
        or with layout "plain" [nsteps][rows][ld] with ld = ldf / ldo (default n)."""
        self._kernel()
        if self.layout == "plain":
            _lib.check(self.lib.hsp2g_run_host_ld(self.h, int(t0), int(nsteps), forcing_ptr, int(ldf or self.n), out_ptr,
                                                  int(ldo or self.n)))
        else:
            _lib.check(self.lib.hsp2g_run_host(self.h, int(t0), int(nsteps), forcing_ptr, out_ptr))

    def run_chunk(self, t0, forcing, out=None):
        """One time chunk in the plain layout, asynchronously: forcing [nsteps][nf][n] of `forcing_dtype` with rows in
        `forcing_rows` order and segments in device order (C-contiguous; pinned memory -- pinned_empty -- for copies that
        overlap the kernels), out [nsteps][ns][n] of `out_dtype` (allocated when None), rows in `save_rows` order.  Nothing
        is re-tiled or copied on the host.  Call sync() before reading `out`."""
        f = forcing
        if self.layout != "plain":
            raise ValueError("run_chunk is the plain-layout call: LandBatch(layout='plain')")
        if f.dtype != self.forcing_dtype or not f.flags.c_contiguous or f.ndim != 3 or f.shape[1:] != (self.nf, self.n):
            raise ValueError("forcing must be C-contiguous %s [nsteps][%d][%d]" % (self.forcing_dtype, self.nf, self.n))
        nsteps = f.shape[0]
        if out is None:
            out = np.empty((nsteps, self.ns, self.n), dtype=self.out_dtype)
        elif out.shape != (nsteps, self.ns, self.n) or out.dtype != self.out_dtype or not out.flags.c_contiguous:
            raise ValueError("out must be C-contiguous %s %r" % (self.out_dtype, (nsteps, self.ns, self.n)))
        self._keep.append((f, out))
        del self._keep[:-4]  # at most two chunks are in flight: older buffers may go
        self._kernel()
        _lib.check(self.lib.hsp2g_run_host_ld(self.h, int(t0), int(nsteps), f.ctypes.data if self.nf and not self.shared else None,
                                              self.n, out.ctypes.data, self.n))
        return out

    def run_device(self, t0, nsteps, d_forcing, d_out, stream=0, ldf=None, ldo=None):
        """Device-resident pass.  layout "tiled": d_forcing [ntiles][nsteps][nf][32], d_out [ntiles][nsteps][ns][32] device
        pointers; "plain": d_forcing [nsteps][nf][ldf], d_out [nsteps][ns][ldo] with ld >= npad (default npad)."""
        self._kernel()
        _lib.check(self.lib.hsp2g_run_device_ld(self.h, int(t0), int(nsteps), d_forcing or None, int(ldf or self.npad), d_out,
                                                int(ldo or self.npad), stream or None))

    def run_device_linked(self, t0, nsteps, sources, src_tiles, d_out, stream=0, ldo=None):
        """Device-resident pass whose forcing rows are read in place from other device buffers (include/hsp2g.h
        hsp2g_run_device_linked).  sources: one (base pointer, interval stride, tile stride) in bytes per forcing row, in
        `self.forcing_rows` order; tile q of this batch reads tile q % src_tiles of the sources."""
        self._kernel(linked=True)
        arr = (_lib.RowSource * max(1, len(sources)))(*[_lib.RowSource(int(b), int(ts), int(tl)) for b, ts, tl in sources])
        _lib.check(self.lib.hsp2g_run_device_linked(self.h, int(t0), int(nsteps), len(sources), arr, int(src_tiles), d_out,
                                                    int(ldo or self.npad), stream or None))

    def set_shared_forcing(self, series, station, mfactor=None):
        """Shared-met mode (include/hsp2g.h hsp2g_set_shared_forcing): `series` [steps][rows][stations] float64 with rows
        in the order of the `forcing_names` given to the constructor, `station` int[n] (which station each segment draws
        on), `mfactor` [rows][n] (default 1.0).  Interval t of segment i then reads series[t][r][station[i]] * mfactor[r][i]
        -- what get_timeseries computes per segment (utilities.py:274-290) -- and run_shared() needs no forcing arrays."""
        series = np.ascontiguousarray(np.asarray(series, dtype=np.float64)[:, self._fperm, :])
        steps, rows, nst = series.shape
        if rows != self.nf:
            raise ValueError("series must have %d rows" % self.nf)
        station = np.asarray(station, dtype=np.int32)
        mf = np.ones((self.nf, self.n)) if mfactor is None else np.asarray(mfactor, dtype=np.float64)[self._fperm, :]
        if station.shape != (self.n,) or mf.shape != (self.nf, self.n):
            raise ValueError("station must be [n], mfactor [rows][n]")
        if self.order is not None:
            station, mf = station[self.order], mf[:, self.order]
        station, mf = np.ascontiguousarray(station), np.ascontiguousarray(mf)
        _lib.check(self.lib.hsp2g_set_shared_forcing(self.h, nst, steps, _ptr(series), _ptr(station, C.POINTER(C.c_int32)),
                                                     _ptr(mf), self.n))
        self.shared_steps = steps
        self.shared = True
        # HSP2G_MET_TABLE=1 and few stations: the kernel fetches an interval's whole station table with one bulk copy instead of
        # per-lane gathers (land_common.cuh met_table_fits; slower on B200, kept as a measured alternative)
        self.met_table = 0 < nst <= 32 and (self.nf * nst) % 2 == 0 and os.environ.get("HSP2G_MET_TABLE", "0")[:1] == "1"

    def run_shared(self, t0=0, steps=None, chunk=None):
        """Advance over [t0, t0+steps) in shared-met mode -> out [steps][ns][n] (caller's row and segment order)."""
        steps = self.shared_steps - t0 if steps is None else steps
        chunk = chunk or steps
        self._kernel()
        if self.layout == "plain":
            dev = np.empty((steps, self.ns, self.n), dtype=self.out_dtype)
            self._keep.append(dev)
            for a in range(0, steps, chunk):
                b = min(steps, a + chunk)
                _lib.check(self.lib.hsp2g_run_host_ld(self.h, int(t0 + a), int(b - a), None, self.n, dev[a:b].ctypes.data, self.n))
            self.sync()
            return self._to_caller(dev)
        out = np.empty((steps, self.ns, self.n), dtype=self.out_dtype)
        tiled = []
        for a in range(0, steps, chunk):
            b = min(steps, a + chunk)
            o = np.empty((self.ntiles, b - a, self.ns, _lib.TILE), dtype=self.out_dtype)
            self._keep.append(o)
            _lib.check(self.lib.hsp2g_run_host(self.h, int(t0 + a), int(b - a), None, o.ctypes.data))
            tiled.append((a, b, o))
        self.sync()
        for a, b, o in tiled:
            u = untile_series(o, self.n)[:, self._sinv, :]
            out[a:b] = u if self._inv is None else u[:, :, self._inv]
        return out

    def _to_device_order(self, forcing):
        """caller's [steps][forcing_names][segments] -> device rows (ascending id) and device segment order; no copy when
        both already agree"""
        f = forcing
        if list(self._fperm) != list(range(self.nf)):
            f = f[:, self._fperm, :]
        if self.order is not None:
            f = f[:, :, self.order]
        return np.ascontiguousarray(f, dtype=self.forcing_dtype)

    def _to_caller(self, dev):
        """device-order series [steps][save_rows][device segments] -> the caller's row and segment order"""
        u = dev
        if list(self._sinv) != list(range(self.ns)):
            u = u[:, self._sinv, :]
        if self._inv is not None:
            u = u[:, :, self._inv]
        return u

    def sync(self):
        _lib.check(self.lib.hsp2g_sync(self.h))
        self._keep = []

    def sync_oldest(self):
        """Wait until the older of the (at most two) chunks in flight through run_chunk / run_host has arrived on the host."""
        _lib.check(self.lib.hsp2g_sync_oldest(self.h))

    def run(self, forcing, chunk=None, out=None):
        """Whole series through host buffers in chunks of `chunk` intervals: forcing [steps][nf][n] -> out
        [steps][ns][n] (plain layouts; re-tiled per chunk on the way in and out).
        out: an array of that shape and the batch's output dtype to fill instead of a new one (saves the first-touch page faults
        of a fresh array, 0.2 s per GB; used as it is when the device's row and segment order are the caller's)."""
        forcing = np.asarray(forcing)
        steps = forcing.shape[0]
        if forcing.shape != (steps, self.nf, self.n):
            raise ValueError("forcing must be [steps][%d][%d]" % (self.nf, self.n))
        chunk = chunk or steps
        if self.layout == "plain":
            # the arrays go to the device as they are: one strided copy per chunk each way, no host re-tiling (a host copy is
            # made only when the caller's row order or a segment permutation differs from the device's); ordinary numpy arrays
            # are page-locked in place for the duration of the call
            f = self._to_device_order(forcing)
            same_order = self._inv is None and list(self._sinv) == list(range(self.ns))
            if out is not None and (out.shape != (steps, self.ns, self.n) or out.dtype != self.out_dtype):
                raise ValueError("out must be [%d][%d][%d] %s" % (steps, self.ns, self.n, self.out_dtype))
            direct = out is not None and same_order and out.flags.c_contiguous
            dev = out if direct else np.empty((steps, self.ns, self.n), dtype=self.out_dtype)
            if not direct:
                _prefault(dev)
            self._keep.append((f, dev))
            self._kernel()
            locked = _PageLocked(self.lib, (f if self.nf else None, dev if self.ns else None))
            try:
                for t0 in range(0, steps, chunk):
                    t1 = min(steps, t0 + chunk)
                    _lib.check(self.lib.hsp2g_run_host_ld(self.h, int(t0), int(t1 - t0), f[t0:t1].ctypes.data if self.nf else None,
                                                          self.n, dev[t0:t1].ctypes.data, self.n))
                self.sync()
            finally:
                try:
                    self.sync()  # nothing may still be copying when the pages are unlocked
                finally:
                    locked.release()
            if out is not None and not direct:
                out[...] = self._to_caller(dev)
                return out
            return self._to_caller(dev)
        forcing = np.asarray(forcing, dtype=self.forcing_dtype)
        out = np.empty((steps, self.ns, self.n), dtype=self.out_dtype)
        tiled = []
        for t0 in range(0, steps, chunk):
            t1 = min(steps, t0 + chunk)
            f = forcing[t0:t1][:, self._fperm, :]
            if self.order is not None:
                f = f[:, :, self.order]
            tiled.append((t0, t1, self.run_host(t0, tile_series(f, self.forcing_dtype))))
        self.sync()
        for t0, t1, o in tiled:
            u = untile_series(o, self.n)[:, self._sinv, :]
            out[t0:t1] = u if self._inv is None else u[:, :, self._inv]
        return out

    def run_periods(self, forcing, period_of_step, periods_per_chunk=32):
        """Period statistics straight from hsp2g_run_host (include/hsp2g.h hsp2g_set_output_periods): forcing [steps][nf][n],
        period_of_step int[steps] (results.period_bins) -> [periods][ns][3 = last, sum, mean][n] float32, the numbers
        IOManager.write_ts computes with pandas for outstep 3 / 4 / 5.  Chunks are cut at period ends."""
        forcing = np.asarray(forcing)
        steps = forcing.shape[0]
        P = np.ascontiguousarray(period_of_step, dtype=np.int32)
        if forcing.shape != (steps, self.nf, self.n) or P.shape != (steps,):
            raise ValueError("forcing must be [steps][%d][%d] and period_of_step [steps]" % (self.nf, self.n))
        if P[0] != 0:
            raise ValueError("period_of_step must start at period 0")
        previous = self.out_dtype
        self.set_output_dtype(np.float32)
        _lib.check(self.lib.hsp2g_set_output_periods(self.h, steps, _ptr(P, C.POINTER(C.c_int32))))
        nper = int(P[-1]) + 1
        starts = np.flatnonzero(np.diff(P, prepend=P[0] - 1))  # first interval of every period present
        cuts = list(starts[::periods_per_chunk]) + [steps]
        if self.layout == "plain":
            f = self._to_device_order(forcing)
            dev = np.empty((nper, self.ns, 3, self.n), dtype=np.float32)
            self._keep.append((f, dev))
            locked = _PageLocked(self.lib, (f if self.nf else None, dev if self.ns else None))
            try:
                self._kernel()
                for a, b in zip(cuts[:-1], cuts[1:]):
                    _lib.check(self.lib.hsp2g_run_host_ld(self.h, int(a), int(b - a), f[a:b].ctypes.data if self.nf else None, self.n,
                                                          dev[int(P[a]):].ctypes.data, self.n))
                self.sync()
            finally:
                try:
                    self.sync()
                finally:
                    locked.release()
                _lib.check(self.lib.hsp2g_set_output_periods(self.h, 0, None))
                self.set_output_dtype(previous)
            u = dev
            if list(self._sinv) != list(range(self.ns)):
                u = u[:, self._sinv]
            return u if self._inv is None else u[:, :, :, self._inv]
        forcing = np.asarray(forcing, dtype=self.forcing_dtype)
        out = np.empty((nper, self.ns, 3, self.n), dtype=np.float32)
        pending = []
        try:
            self._kernel()
            for a, b in zip(cuts[:-1], cuts[1:]):
                f = forcing[a:b][:, self._fperm, :]
                if self.order is not None:
                    f = f[:, :, self.order]
                ft = tile_series(f, self.forcing_dtype)
                nb = int(P[b - 1] - P[a]) + 1
                o = np.empty((self.ntiles, nb, self.ns, 3, _lib.TILE), dtype=np.float32)
                self._keep.append((ft, o))
                _lib.check(self.lib.hsp2g_run_host(self.h, int(a), int(b - a), ft.ctypes.data, o.ctypes.data))
                pending.append((int(P[a]), nb, o))
            self.sync()
        finally:
            self.sync()  # nothing of this call may still be in flight when periods and dtype are restored
            _lib.check(self.lib.hsp2g_set_output_periods(self.h, 0, None))
            self.set_output_dtype(previous)
        for p0, nb, o in pending:
            u = o.transpose(1, 2, 3, 0, 4).reshape(nb, self.ns, 3, self.ntiles * _lib.TILE)[:, :, :, :self.n][:, self._sinv]
            out[p0:p0 + nb] = u if self._inv is None else u[:, :, :, self._inv]
        return out

    # ------------------------------------------------------------------------------------------------
    def errors(self):
        """{error id: int64[n]} for the counters of the active sections."""
        E = _lib.index("errors")
        buf = np.zeros((len(E), self.n), dtype=np.int64)
        _lib.check(self.lib.hsp2g_get_errors(self.h, _ptr(buf, C.POINTER(C.c_int64)), self.n))
        if self._inv is not None:
            buf = buf[:, self._inv]
        return {nme: buf[i] for nme, i in E.items()}

    def errors_for(self, section, seg=0):
        e = self.errors()
        return np.array([e[k][seg] for k in ERRORS_OF[section]], dtype=np.int64)

    def get_state(self):
        S = _lib.index("states")
        buf = np.zeros((len(S), self.n), dtype=np.float64)
        _lib.check(self.lib.hsp2g_get_state(self.h, _ptr(buf), self.n))
        return buf if self._inv is None else np.ascontiguousarray(buf[:, self._inv])

    def set_state(self, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        if self.order is not None:
            state = np.ascontiguousarray(state[:, self.order])
        _lib.check(self.lib.hsp2g_set_state(self.h, _ptr(state), self.n))

    def set_output_dtype(self, dtype):
        """np.float64 (default) or np.float32: element type of the saved series the kernels write.  float32 is what
        the reference's save_timeseries hands the IOManager (utilities.py:313)."""
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("output dtype must be float64 or float32")
        _lib.check(self.lib.hsp2g_set_output_dtype(self.h, _lib.F32 if dtype == np.float32 else _lib.F64))
        self.out_dtype = dtype

    def set_schedule(self, sub_steps):
        """Intervals per (tile, sub-chunk) task of one launch (include/hsp2g.h hsp2g_set_schedule); 0 = whole chunk."""
        _lib.check(self.lib.hsp2g_set_schedule(self.h, int(sub_steps)))

    def set_launch_overlap(self, on=True):
        """Let consecutive run_device calls on one stream overlap at their tails (include/hsp2g.h hsp2g_set_launch_overlap: the
        buffers of a call must be ready when the previous call is enqueued, outputs distinct from the previous call's)."""
        _lib.check(self.lib.hsp2g_set_launch_overlap(self.h, 1 if on else 0))

    def chained_launches(self):
        return int(self.lib.hsp2g_chained_launches(self.h))

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


class GroupedLandBatch:
    """An ensemble whose segments carry SEVERAL option words, run as one group per word: the segments are sorted by word
    (divergence-aware sorting), every group is a LandBatch of its own -- so it gets the kernel built for ITS word, with every
    option test folded away (jit.py) -- and the groups are launched together over column ranges of one series buffer
    (hsp2g_run_device_group), sharing the GPU.  A mixed ensemble otherwise runs the one kernel that tests the options at run
    time: larger code, more registers, fewer resident warps (profiles/README.md: 250k full-chain segments with 48 words).

    Device columns: group g owns [first_column[g], first_column[g] + npad_g); `column_of[i]` is the device column of the caller's
    segment i.  run() takes and returns arrays in the caller's order; run_device() works on a buffer of `width` columns."""

    def __init__(self, store, calendar, forcing_names, save, device=0, layout="plain", jit=None, keys=(), min_group=1, **kw):
        self.store, self.cal, self.n = store, calendar, store.n
        words, inv = np.unique(store.flags, return_inverse=True)
        keys = tuple(np.asarray(k) for k in keys)
        self.groups, self.first_column, self.members = [], [], []
        col = 0
        try:
            for w in range(len(words)):
                idx = np.flatnonzero(inv == w)
                if keys:  # climate / caller keys inside a group (most significant first)
                    idx = idx[np.lexsort([k[idx] for k in reversed(keys)])]
                b = LandBatch(store.subset(idx), calendar, forcing_names, save, device=device, layout=layout, jit=jit, order=None, **kw)
                self.groups.append(b)
                self.members.append(idx)
                self.first_column.append(col)
                col += b.npad
        except Exception:
            self.close()
            raise
        self.width = col
        self.column_of = np.empty(self.n, dtype=np.int64)
        for c0, idx in zip(self.first_column, self.members):
            self.column_of[idx] = c0 + np.arange(len(idx))
        g0 = self.groups[0]
        self.lib, self.layout, self.device = g0.lib, layout, device
        self.nf, self.ns, self.forcing_rows, self.save_rows, self.save = g0.nf, g0.ns, g0.forcing_rows, g0.save_rows, g0.save
        self.forcing_names = g0.forcing_names
        self._handles = (_lib._h * len(self.groups))(*[b.h for b in self.groups])
        self._cols = np.ascontiguousarray(self.first_column, dtype=np.int64)
        if any(b.jit for b in self.groups):  # build the groups' kernels side by side, not one after the other at first launch
            from . import jit as _jit

            try:
                _jit.prebuild([_jit.spec_of(b) for b in self.groups], _lib.build_kind())
            except Exception:
                pass  # a group whose kernel cannot be built runs the library's general kernel (LandBatch._kernel records why)

    def run_device(self, t0, nsteps, d_forcing, d_out, stream=0, ldf=None, ldo=None):
        """d_forcing / d_out: device buffers of `width` columns -- PLAIN [nsteps][rows][ld] (ld default width) or TILED
        [width/32][nsteps][rows][32] -- with group g's segments in columns first_column[g].. in `members[g]` order."""
        for b in self.groups:
            b._kernel()
        _lib.check(self.lib.hsp2g_run_device_group(len(self.groups), self._handles, int(t0), int(nsteps), d_forcing or None,
                                                   int(ldf or self.width), d_out, int(ldo or self.width),
                                                   self._cols.ctypes.data_as(C.POINTER(C.c_int64)), stream or None))

    def run(self, forcing, chunk=None):
        """forcing [steps][nf][n] in the caller's segment order -> out [steps][ns][n]; host path, one group after the other."""
        forcing = np.asarray(forcing)
        out = np.empty((forcing.shape[0], self.ns, self.n), dtype=self.groups[0].out_dtype)
        for b, idx in zip(self.groups, self.members):
            out[:, :, idx] = b.run(np.ascontiguousarray(forcing[:, :, idx]), chunk=chunk)
        return out

    def _gather(self, fn, dtype):
        parts = [fn(b) for b in self.groups]
        out = np.empty(parts[0].shape[:-1] + (self.n,), dtype=dtype)
        for p, idx in zip(parts, self.members):
            out[..., idx] = p
        return out

    def get_state(self):
        return self._gather(lambda b: b.get_state(), np.float64)

    def errors(self):
        per = [b.errors() for b in self.groups]
        out = {k: np.empty(self.n, dtype=np.int64) for k in per[0]}
        for e, idx in zip(per, self.members):
            for k, v in e.items():
                out[k][idx] = v
        return out

    def set_output_dtype(self, dtype):
        for b in self.groups:
            b.set_output_dtype(dtype)

    def set_schedule(self, sub_steps):
        for b in self.groups:
            b.set_schedule(sub_steps)

    def sync(self):
        for b in self.groups:
            b.sync()

    def kernel_names(self):
        return [b.kernel_name() for b in self.groups]

    def close(self):
        for b in getattr(self, "groups", ()):
            b.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class QualBatch(LandBatch):
    """PQUAL / IQUAL of many segments: one device row per (segment, constituent) (qualstore.qual_store).

    `uci_list[i]` is segment i's PQUAL / IQUAL table set.  `forcing_names` are the input series the rows read (qualstore.
    PQUAL_INPUTS / IQUAL_INPUTS); `gather()` spreads per-SEGMENT series -- the land batch's SURO, IFWO, ... outputs and the
    met / lateral inputs -- over the rows, `run()` then works as for any batch.  Output row j belongs to `self.rows[j]` =
    (segment index, constituent number); in the reference's ts that series is '<P|I>QUAL<k>_<NAME>'."""

    def __init__(self, operation, uci_list, calendar, forcing_names, save="all", device=0, timing=False, order="flags",
                 land=None, layout="plain", jit=None, chain=None):
        """chain (with land=...): "linked" -- the rows read the land batch's device buffers IN PLACE (run_device_linked_from; no
        scratch buffer, no gather launch).  The device rows are constituent-major over the land batch's tiles: device row
        k*land.npad + p is constituent slot k of the land batch's device segment p (`self.rows[j]` = (caller's segment, constituent
        number), segment -1 for the lanes that pad the land batch's last tile), so a tile of 32 rows reads exactly one tile of
        every land series.  Needs the same number of constituents in every segment and no deposition input series.  "gather" --
        the rows keep an order of their own and a device gather fills a scratch buffer (run_device_from).  None = "linked" where
        possible.
        order="flags" (default) puts rows with equal option words next to each other on the device: the constituents of a
        segment differ in kind (sediment-associated, overland-flow-associated, subsurface concentrations ...), segment-major rows
        would make every warp execute every kind.  Invisible to run() / gather(); run_device buffers are in device order.
        land: the LandBatch whose device buffers will feed this batch (attach() is called): within equal option words the rows
        then follow the land batch's device order, so that 32 consecutive rows gather 32 consecutive cells of a land tile."""
        from .qualstore import nquals, qual_store

        self.chain = None
        if land is not None and chain != "gather":
            counts = {nquals(u) for u in uci_list}
            if len(counts) == 1 and len(uci_list) == land.n:
                # the land batch's device order, its last tile padded with copies of its last segment (as its own lanes are)
                pos = np.arange(land.npad)
                pos[land.n:] = land.n - 1
                seg = pos if land.order is None else np.asarray(land.order)[pos]
                nq = counts.pop()
                ext = qual_store(operation, [uci_list[i] for i in seg], calendar)
                if not any(ext.dep_series) and nq > 0:
                    self.chain = "linked"
                    store = ext
                    store.rows = [(int(seg[p]) if p < land.n else -1, k) for p, k in store.rows]
                    # qual_store's rows are segment-major (p, k): device row k*npad + p holds its row p*nq + k
                    order = (np.arange(land.npad)[None, :] * nq + np.arange(nq)[:, None]).reshape(-1)
            if self.chain is None and chain == "linked":
                raise NotImplementedError("chain='linked' needs one constituent count for all segments, one table set per land "
                                          "segment and no deposition input series")
        if self.chain is None:
            store = qual_store(operation, uci_list, calendar)
            if land is not None:
                self.chain = "gather"
        self.dep_keys = None
        if any(store.dep_series):
            # Some constituent takes its atmospheric deposition from an input series (PQADFG / IQADFG = -1, PQUAL.py:94-103,
            # IQUAL.py:88-94).  The deposition flux / concentration then are forcing rows of the batch (QADFX, QADCN), one column
            # per (segment, constituent) row: gather() fills a row's column from the caller's series where the row's flag is
            # -1 and, elsewhere, with the very series the reference's wrapper builds for it (its monthly table interpolated to
            # the day, or zeros) -- what the kernel would have computed itself.
            if land is not None:
                raise NotImplementedError("deposition input series with the device-resident chain (land=...): feed the rows "
                                          "through gather() / run()")
            forcing_names = [n for n in forcing_names if n not in ("QADFX", "QADCN")] + ["QADFX", "QADCN"]
            sec, ad = ("IQUAL", "IQAD") if operation == "IMPLND" else ("PQUAL", "PQAD")
            self.dep_keys, self._dep_base = [], np.zeros((calendar.steps, 2, store.n))
            for j, (i, k) in enumerate(store.rows):
                uci = uci_list[i]
                keys = {}
                u = uci.get("FLAGS")
                for c, nm in enumerate(("FX", "CN")):
                    val = u[ad + "FG" + str(2 * k - 1 + c)] if u is not None else 0
                    if val > 0:
                        self._dep_base[:, c, j] = calendar.initm(uci, val, "%s%d_MONTHLY/%s%s" % (sec, k, ad, nm), 0.0)[:calendar.steps]
                    elif val == -1:
                        keys[nm] = "%s%s%d" % (ad, nm, k)  # the caller's series: deposition[(segment index, this key)]
                self.dep_keys.append(keys)
        self.rows = store.rows
        self.row_segment = np.array([i for i, _k in store.rows], dtype=np.int64)
        if self.chain == "gather" and isinstance(order, str) and order == "flags":
            pos = self.row_segment if land._inv is None else np.asarray(land._inv)[self.row_segment]
            order = store.divergence_order(keys=(pos,))
        super().__init__(store, calendar, forcing_names, save, device=device, timing=timing, order=order,
                         layout="tiled" if self.chain == "gather" else layout, jit=jit)
        if land is not None:
            self.attach(land)

    def gather(self, series, deposition=None):
        """{name: [steps][n_segments] or [steps]} -> forcing [steps][nf][rows] in `forcing_names` order.
        deposition: {(segment index, 'PQADFX<k>' | 'PQADCN<k>' | 'IQADFX<k>' | 'IQADCN<k>'): float64[steps]} for the constituents
        whose deposition flag is -1 (`dep_keys[row]` names what each row needs)."""
        steps = len(next(iter(series.values())))
        out = np.empty((steps, len(self.forcing_names), self.n))
        for r, nme in enumerate(self.forcing_names):
            if self.dep_keys is not None and nme in ("QADFX", "QADCN"):
                c = 0 if nme == "QADFX" else 1
                out[:, r, :] = self._dep_base[:steps, c, :]
                for j, keys in enumerate(self.dep_keys):
                    key = keys.get("FX" if c == 0 else "CN")
                    if key is not None:
                        if deposition is None or (self.rows[j][0], key) not in deposition:
                            raise KeyError("deposition series %r of segment %d is an input (flag -1) and was not given" % (key, self.rows[j][0]))
                        out[:, r, j] = np.asarray(deposition[(self.rows[j][0], key)], dtype=np.float64)[:steps]
                continue
            a = np.asarray(series[nme], dtype=np.float64)
            out[:, r, :] = a[:steps, None] if a.ndim == 1 else a[:steps][:, self.row_segment]
        return out

    # -- device-resident chain: land batch -> gather -> constituents (no host round trip of the flow series) ------------------
    OUTPUT_OF = {"PERLND": {"SURO": "PWATER_SURO", "IFWO": "PWATER_IFWO", "AGWO": "PWATER_AGWO", "PERO": "PWATER_PERO",
                            "WSSD": "SEDMNT_WSSD", "SCRSD": "SEDMNT_SCRSD"},
                 "IMPLND": {"SURO": "IWATER_SURO", "SOSLD": "SOLIDS_SOSLD"}}

    def attach(self, land):
        """Take this batch's input series from `land`'s device buffers: a series the land sections produce comes from
        land's output rows (it must be in land's SAVE set, float64), anything else from land's forcing rows."""
        if land.out_dtype != np.float64:
            raise ValueError("the land batch must write float64 outputs to feed constituents")
        if self.chain == "linked":
            self._link = []  # per forcing row of this batch: ('out' | 'forc', the land batch's row)
            for nme in self.forcing_rows:
                full = self.OUTPUT_OF[self.store.operation].get(nme)
                if full is not None and full in land.save_rows:
                    self._link.append(("out", land.save_rows.index(full)))
                elif nme in land.forcing_rows and full is None:
                    self._link.append(("forc", land.forcing_rows.index(nme)))
                else:
                    raise KeyError("%s: neither saved by the land batch (%s) nor one of its forcing rows" % (nme, full))
            self._land = land
            return self
        if land.layout != "tiled" or self.layout != "tiled":
            raise ValueError("the gathering chain works between tile-major buffers: LandBatch(layout='tiled')")
        rows = self.row_segment if self.order is None else self.row_segment[self.order]  # device row -> caller's segment
        seg = rows if land._inv is None else np.asarray(land._inv)[rows]  # -> the land batch's device position
        seg = np.ascontiguousarray(seg, dtype=np.int64)
        self._gather = _lib._h()
        _lib.check(self.lib.hsp2g_gather_create(self.device, land.npad, self.n, _ptr(seg, C.POINTER(C.c_int64)),
                                                C.byref(self._gather)))
        from_out, from_forc = [], []
        for r, nme in enumerate(self.forcing_rows):
            full = self.OUTPUT_OF[self.store.operation].get(nme)
            if full is not None and full in land.save_rows:
                from_out.append((land.save_rows.index(full), r))
            elif nme in land.forcing_rows and full is None:
                from_forc.append((land.forcing_rows.index(nme), r))
            else:
                raise KeyError("%s: neither saved by the land batch (%s) nor one of its forcing rows" % (nme, full))
        self._plan = (land.ns, from_out, land.nf, from_forc)
        self._land = land
        return self

    def run_device_from(self, t0, nsteps, d_land_forcing, d_land_out, d_scratch, d_out, stream=0):
        """Gather the inputs of intervals [t0, t0+nsteps) from the land batch's device buffers of the same chunk into
        `d_scratch` ([ntiles][nsteps][nf][32] float64) and run the constituents into `d_out`.  With an explicit `stream` the
        caller orders this after the land launch (same stream); without one the land batch is synchronised first and the
        gather and the constituents run on this batch's own stream."""
        ns, from_out, nf, from_forc = self._plan
        if not stream:
            self._land.sync()
            stream = self.lib.hsp2g_batch_stream(self.h)
        i32 = C.POINTER(C.c_int32)
        for src, rows, pairs in ((d_land_out, ns, from_out), (d_land_forcing, nf, from_forc)):
            if not pairs:
                continue
            a = np.array([p[0] for p in pairs], dtype=np.int32)
            b = np.array([p[1] for p in pairs], dtype=np.int32)
            _lib.check(self.lib.hsp2g_gather_run(self._gather, src, rows, d_scratch, self.nf, len(pairs), _ptr(a, i32),
                                                 _ptr(b, i32), int(nsteps), stream or None))
        self.run_device(t0, nsteps, d_scratch, d_out, stream)

    def run_device_linked_from(self, t0, nsteps, d_land_forcing, d_land_out, d_out, stream=0, land_ldf=None, land_ldo=None,
                               ldo=None):
        """chain="linked": run the constituents of intervals [t0, t0+nsteps) straight from the land batch's device buffers of the
        same chunk (its forcing buffer, float64, and the output buffer it wrote; the land batch's layout, row lengths land_ldf /
        land_ldo when plain) into d_out (this batch's layout).  Stream ordering as for run_device_from."""
        land = self._land
        if land.forcing_dtype != np.float64 and any(kind == "forc" for kind, _r in self._link):
            raise ValueError("linked rows are float64: the land batch's forcings must be float64")
        if not stream:
            land.sync()
            stream = self.lib.hsp2g_batch_stream(self.h)
        src = []
        for kind, row in self._link:
            buf, rows, ld = (d_land_out, land.ns, land_ldo) if kind == "out" else (d_land_forcing, land.nf, land_ldf)
            if land.layout == "tiled":
                src.append((buf + row * 256, rows * 256, int(nsteps) * rows * 256))
            else:
                ld = int(ld or land.npad)
                src.append((buf + row * ld * 8, rows * ld * 8, 256))
        self.run_device_linked(t0, nsteps, src, land.ntiles, d_out, stream, ldo)

    def close(self):
        g = getattr(self, "_gather", None)
        if g is not None and g.value:
            self.lib.hsp2g_gather_destroy(g)
            self._gather = None
        super().close()

    def segment_errors(self, n_segments):
        """int64[2][n_segments]: the reference's error vector of each segment (errors[0] = constituents that take
        atmospheric deposition without being associated with overland flow, PQUAL.py:171-174)."""
        e = np.zeros((2, n_segments), dtype=np.int64)
        real = self.row_segment >= 0  # chain="linked": the rows of the land batch's padding lanes belong to no segment
        np.add.at(e[0], self.row_segment[real], self.store.dep_without_qsofg[real])
        return e


def tile_series(x, dtype=np.float64):
    """[steps][rows][n] -> tile-major [ntiles][steps][rows][32] (include/hsp2g.h); lanes past n copy the last segment."""
    x = np.asarray(x, dtype=dtype)
    steps, rows, n = x.shape
    T = _lib.TILE
    nt = (n + T - 1) // T
    if nt * T != n:
        x = np.concatenate([x, np.repeat(x[:, :, -1:], nt * T - n, axis=2)], axis=2)
    return np.ascontiguousarray(x.reshape(steps, rows, nt, T).transpose(2, 0, 1, 3))


def untile_series(y, n):
    """tile-major [ntiles][steps][rows][32] -> [steps][rows][n]."""
    nt, steps, rows, T = y.shape
    return y.transpose(1, 2, 0, 3).reshape(steps, rows, nt * T)[:, :, :n]


def pack_forcings(series, names, n):
    """{name: float64[steps] or [steps][n]} -> [steps][len(names)][n] (shared series are broadcast)."""
    steps = len(next(iter(series.values())))
    out = np.empty((steps, len(names), n), dtype=np.float64)
    for r, nme in enumerate(names):
        a = np.asarray(series[nme], dtype=np.float64)
        out[:, r, :] = a[:steps, None] if a.ndim == 1 else a[:steps]
    return out
