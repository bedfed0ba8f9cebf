"""ctypes binding of libhsp2g.so (include/hsp2g.h).  There is no CPU fallback: if the shared object is missing or no
CUDA device is present, the engine fails loudly.

The id spaces (parameters, monthly tables, flag bits, states, forcings, outputs, error counters) are read back from the
library itself (hsp2g_names), so host code can never drift from the kernels.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HSP2G_LIB selects another build of the SAME library (kernel tuning experiments, hspsquared_b200/build.py); whatever the
# path, load() accepts only a library that says it is the CUDA build (hsp2g_names("build")), so no host-compiled stand-in
# can ever serve the product path.  The CPU test suite's double is admitted by tests/conftest.py alone, through ALLOW_BUILDS.
LIB_PATH = os.environ.get("HSP2G_LIB") or os.path.join(_HERE, "libhsp2g.so")
CUDA_BUILD = "cuda/sm_100a"
ALLOW_BUILDS = {CUDA_BUILD}

PERLND, IMPLND = 0, 1
ABI_VERSION = 3
F64, F32 = 0, 1
TILED, PLAIN = 0, 1
HOST_WRITE_COMBINED, HOST_PORTABLE = 1, 2
TILE = 32  # segments per tile of the series layout [tile][interval][row][32] (include/hsp2g.h)
ACT_BITS = {"ATEMP": 1, "SNOW": 2, "PWATER": 4, "SEDMNT": 8, "PSTEMP": 16, "PWTGAS": 32, "IWATER": 64, "SOLIDS": 128,
            "IWTGAS": 256, "PQUAL": 512, "IQUAL": 1024}
OPT_SED_RAIN_IS_PREC = 1
OPT_METRIC = 2
CAL_HRFG, CAL_DAYFG, CAL_HR1FG, CAL_HR6IND, CAL_SEASON, CAL_NEWDAY = 1, 2, 4, 8, 16, 32

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)
_h = C.c_void_p



class RowSource(C.Structure):
    """hsp2g_row_source (include/hsp2g.h): where one linked forcing row lives, strides in bytes"""
    _fields_ = [("base", C.c_void_p), ("interval_stride", C.c_int64), ("tile_stride", C.c_int64)]


_PROTOS = {
    "hsp2g_abi_version": (C.c_int, []),
    "hsp2g_last_error": (C.c_char_p, []),
    "hsp2g_names": (C.c_char_p, [C.c_char_p]),
    "hsp2g_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "hsp2g_batch_create": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_uint32, C.c_double, C.POINTER(_h)]),
    "hsp2g_batch_destroy": (C.c_int, [_h]),
    "hsp2g_batch_npad": (C.c_int64, [_h]),
    "hsp2g_set_params": (C.c_int, [_h, _dp, C.c_int64]),
    "hsp2g_set_flags": (C.c_int, [_h, _u64p]),
    "hsp2g_set_monthly": (C.c_int, [_h, C.c_int, _dp, C.c_int64]),
    "hsp2g_set_state": (C.c_int, [_h, _dp, C.c_int64]),
    "hsp2g_get_state": (C.c_int, [_h, _dp, C.c_int64]),
    "hsp2g_set_calendar": (C.c_int, [_h, C.c_int64, _u8p, _u8p, _u8p, _dp, _dp, _dp]),
    "hsp2g_set_forcing_layout": (C.c_int, [_h, C.c_int, _i32p, _dp, C.c_uint32]),
    "hsp2g_set_shared_forcing": (C.c_int, [_h, C.c_int32, C.c_int64, _dp, _i32p, _dp, C.c_int64]),
    "hsp2g_set_save": (C.c_int, [_h, C.c_int, _i32p]),
    "hsp2g_run_device": (C.c_int, [_h, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hsp2g_run_host": (C.c_int, [_h, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "hsp2g_run_device_ld": (C.c_int, [_h, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "hsp2g_run_host_ld": (C.c_int, [_h, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "hsp2g_run_device_linked": (C.c_int, [_h, C.c_int64, C.c_int32, C.c_int32, C.POINTER(RowSource), C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_void_p]),
    "hsp2g_run_device_group": (C.c_int, [C.c_int32, C.POINTER(_h), C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         _i64p, C.c_void_p]),
    "hsp2g_set_series_layout": (C.c_int, [_h, C.c_int]),
    "hsp2g_set_forcing_dtype": (C.c_int, [_h, C.c_int]),
    "hsp2g_set_kernel_image": (C.c_int, [_h, C.c_void_p, C.c_size_t, C.c_char_p, C.c_char_p]),
    "hsp2g_aggregate_f32": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, _i32p, C.c_int32, C.c_void_p, C.c_void_p]),
    "hsp2g_set_output_periods": (C.c_int, [_h, C.c_int64, _i32p]),
    "hsp2g_batch_stream": (C.c_void_p, [_h]),
    "hsp2g_gather_create": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _i64p, C.POINTER(_h)]),
    "hsp2g_gather_run": (C.c_int, [_h, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, _i32p, _i32p, C.c_int32, C.c_void_p]),
    "hsp2g_gather_destroy": (None, [_h]),
    "hsp2g_pack_series": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "hsp2g_unpack_series": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]),
    "hsp2g_sync": (C.c_int, [_h]),
    "hsp2g_sync_oldest": (C.c_int, [_h]),
    "hsp2g_get_errors": (C.c_int, [_h, _i64p, C.c_int64]),
    "hsp2g_reset_errors": (C.c_int, [_h]),
    "hsp2g_set_output_dtype": (C.c_int, [_h, C.c_int]),
    "hsp2g_set_schedule": (C.c_int, [_h, C.c_int32]),
    "hsp2g_set_launch_overlap": (C.c_int, [_h, C.c_int]),
    "hsp2g_chained_launches": (C.c_int64, [_h]),
    "hsp2g_set_timing": (C.c_int, [_h, C.c_int]),
    "hsp2g_kernel_stats": (C.c_int, [_h, _i64p, _dp]),
    "hsp2g_kernel_name": (C.c_char_p, [_h]),
    "hsp2g_math_probe": (C.c_int, [C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp]),
    "hsp2g_host_alloc": (C.c_void_p, [C.c_size_t]),
    "hsp2g_host_alloc_flags": (C.c_void_p, [C.c_size_t, C.c_uint]),
    "hsp2g_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "hsp2g_host_unregister": (C.c_int, [C.c_void_p]),
    "hsp2g_host_free": (C.c_int, [C.c_void_p]),
    "hsp2g_device_alloc": (C.c_void_p, [C.c_int, C.c_size_t]),
    "hsp2g_device_free": (C.c_int, [C.c_int, C.c_void_p]),
    "hsp2g_memcpy_h2d": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "hsp2g_memcpy_d2h": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
}
EXPORTS = tuple(_PROTOS)

_lib = None
_names = {}


class Hsp2gError(RuntimeError):
    pass


def load():
    """Load libhsp2g.so; raises if it has not been built (python -m hspsquared_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Hsp2gError("%s is missing: build it with `python -m hspsquared_b200.build` (nvcc, sm_100a). "
                         "There is no CPU fallback for the land-segment hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.hsp2g_abi_version() != ABI_VERSION:
        raise Hsp2gError("libhsp2g ABI version mismatch")
    kind = lib.hsp2g_names(b"build").decode()
    if kind not in ALLOW_BUILDS:
        raise Hsp2gError("%s is a %r build of libhsp2g; the land-segment path runs on the CUDA build (%s) only -- there is no "
                         "CPU fallback" % (LIB_PATH, kind, CUDA_BUILD))
    _lib = lib
    return lib


def build_kind():
    """'cuda/sm_100a' for the product library (load() admits nothing else outside the CPU test suite)."""
    return load().hsp2g_names(b"build").decode()


def check(rc):
    if rc != 0:
        raise Hsp2gError(load().hsp2g_last_error().decode())


def names(table):
    """Ordered id list of one identifier space, straight from the library."""
    if table not in _names:
        s = load().hsp2g_names(table.encode()).decode()
        _names[table] = tuple(x for x in s.split(",") if x)
    return _names[table]


def index(table):
    return {n: i for i, n in enumerate(names(table))}


def device_count():
    n = C.c_int(0)
    rc = load().hsp2g_device_count(C.byref(n))
    return n.value if rc == 0 else 0
