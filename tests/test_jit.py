"""Kernels built for one batch (hspsquared_b200/jit.py) through the host-compiled TEST DOUBLE: the generated translation unit
is the same text that nvcc compiles for the GPU; here g++ builds it (tests/emu), the double loads it through the same C-ABI call
(hsp2g_set_kernel_image) and applies the same "does the image still describe the batch" check before every launch."""
import ctypes as C

import numpy as np
import pytest

from hspsquared_b200 import _lib, jit, synth, test10
from hspsquared_b200.calendar import Calendar
from hspsquared_b200.engine import LandBatch

from . import parity


@pytest.fixture(autouse=True)
def _use_emu(emu_library):
    yield


CAL = Calendar("1976-02-10", "1976-02-20", 60, "us")


def _perlnd(n=40):
    ens = synth.Ensemble("PERLND", n, seed=3, sections=("SNOW", "PWATER"))
    return ens, synth.forcing_chunk(ens, CAL, 0, CAL.steps)


def test_generated_source_names_what_it_fixes():
    ens, _f = _perlnd()
    with LandBatch(ens.store(CAL), CAL, list(synth.PERLND_FORCINGS), ["PWATER_SURO", "SNOW_PACKF"], jit=False) as b:
        spec = jit.spec_of(b)
    assert spec["word"] == int(ens.store(CAL).flags[0]) and spec["plain"] and not spec["f32in"] and spec["act"] == 6
    text = jit.source(spec, 128)
    assert "StaticFlags<%#xull, 0xffffffffffffffffull, 0>" % spec["word"] in text and "StaticLay<true, false, -1, false>" in text
    assert "perlnd_body<ACT, HOURLY, IO, SHARED, FW, LAY>" in text and "__launch_bounds__(LAND_BLOCK, 8)" in text
    # without a common option word the bits are read per segment (mask 0); the batch-wide HSP2G_OPT_* bits stay fixed
    assert "StaticFlags<0x0ull, 0x0ull, 2>" in jit.source(dict(spec, word=None, opts=2), 168) and "__launch_bounds__(LAND_BLOCK, 6)" in jit.source(spec, 168)


def test_dtype_and_layout_changes_rebuild_and_stale_images_are_refused():
    ens, f = _perlnd()
    save = ["SNOW_PACKF", "PWATER_SURO", "PWATER_AGWS"]
    with LandBatch(ens.store(CAL), CAL, list(synth.PERLND_FORCINGS), save, jit=False) as b:
        ref = b.run(f)
    with LandBatch(ens.store(CAL), CAL, list(synth.PERLND_FORCINGS), save, jit=True) as b:
        a = b.run(f[:100])
        k64 = b.kernel_name()
        dev_rows = [b.save.index(r) for r in b.save_rows]  # the raw calls below speak the device's row order
        assert ",jit>" in k64 and b.jit_error is None
        # the library itself refuses the image once the batch no longer is what the image was built for ...
        _lib.check(b.lib.hsp2g_set_output_dtype(b.h, _lib.F32))
        o = np.empty((10, b.ns, b.n), dtype=np.float32)
        _lib.check(b.lib.hsp2g_run_host_ld(b.h, 100, 10, np.ascontiguousarray(f[100:110]).ctypes.data, b.n, o.ctypes.data, b.n))
        b.sync()
        assert "jit" not in b.kernel_name(), "a float64 image ran on a float32 batch"
        # ... and the engine builds the matching one when the change goes through it
        b.set_output_dtype(np.float32)
        c = b.run_chunk(110, np.ascontiguousarray(f[110:]))
        b.sync()
        assert ",f32,plain" in b.kernel_name() and b.kernel_name() != k64
    assert np.array_equal(a, ref[:100], equal_nan=True)
    with np.errstate(over="ignore"):
        assert np.array_equal(o, ref[100:110][:, dev_rows].astype(np.float32), equal_nan=True)
        assert np.array_equal(c, ref[110:][:, dev_rows].astype(np.float32), equal_nan=True)


@pytest.mark.parametrize("layout", ["plain", "tiled"])
def test_float32_forcings_are_widened_exactly(layout):
    ens, f = _perlnd(33)  # a ragged last tile
    f32 = f.astype(np.float32)
    with LandBatch(ens.store(CAL), CAL, list(synth.PERLND_FORCINGS), "all", jit=False, layout="tiled") as b:
        ref = b.run(f32.astype(np.float64), chunk=77)
    for use_jit in (False, True):
        with LandBatch(ens.store(CAL), CAL, list(synth.PERLND_FORCINGS), "all", jit=use_jit, layout=layout) as b:
            b.set_forcing_dtype(np.float32)
            got = b.run(f32, chunk=77)
            assert (",jit>" in b.kernel_name()) == use_jit
        assert np.array_equal(got, ref, equal_nan=True), (layout, use_jit)


def test_metric_units_are_compiled_into_the_batch_kernel():
    parity.check_metric_batch_kernels()


def test_shared_met_station_table_mode():
    """HSP2G_MET_TABLE=1: the interval's station table by one bulk copy (<= 32 stations; more fall back to the gathers)."""
    parity.check_shared_met_mode(n=75, jit=True, table=True)
    parity.check_shared_met_mode(n=75, jit=True, nst=37, table=True)


def test_shared_met_implnd_and_constituent_kernels():
    parity.check_shared_met_mode(n=75, jit=True)
    ens = synth.Ensemble("IMPLND", 40, seed=4, sections=("SNOW", "IWATER", "SOLIDS"))
    f = synth.forcing_chunk(ens, CAL, 0, CAL.steps)
    outs = []
    for use_jit in (False, True):
        with LandBatch(ens.store(CAL), CAL, list(synth.PERLND_FORCINGS), "all", jit=use_jit) as b:
            outs.append(b.run(f, chunk=100))
            assert ("implnd_kernel<SNOW|IWATER|SOLIDS" in b.kernel_name() and ",jit>" in b.kernel_name()) == use_jit
    assert np.array_equal(outs[0], outs[1], equal_nan=True)
    from hspsquared_b200.engine import QualBatch
    from hspsquared_b200.qualstore import PQUAL_INPUTS

    rng = np.random.default_rng(5)
    series = {k: rng.random((CAL.steps, 6)) * (rng.random((CAL.steps, 6)) > 0.5) for k in PQUAL_INPUTS[:7]}
    outs = []
    for use_jit in (False, True):
        with QualBatch("PERLND", [test10.pqual_tables()] * 6, CAL, list(PQUAL_INPUTS[:7]), "all", jit=use_jit) as q:
            outs.append(q.run(q.gather(series), chunk=90))
            assert ("qual_kernel<PQUAL" in q.kernel_name() and ",jit>" in q.kernel_name()) == use_jit
    assert np.array_equal(outs[0], outs[1], equal_nan=True)
