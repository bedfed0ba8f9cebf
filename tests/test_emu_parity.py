"""CPU suite: the real host stack (segment store, calendar, C ABI, chunking, drop-in activities) and the kernels'
control flow, executed through the host-compiled TEST DOUBLE of libhsp2g (tests/emu/README.md) and compared with the
CPU oracle BIT FOR BIT (same glibc libm, no FMA contraction).  The CUDA library itself is exercised by the `-m gpu`
suite (tests/test_gpu_parity.py) through the same harness."""
import math

import numpy as np
import pytest

from oracle import cases as CASES

from . import parity

ALL = CASES.build_cases()


@pytest.fixture(scope="module", autouse=True)
def _use_emu(emu_library):
    yield


@pytest.mark.parametrize("name", sorted(ALL))
def test_fused_bit_exact(name):
    case = ALL[name]
    ref, ref_err = parity.run_oracle(case)
    got, got_err, _ = parity.run_fused(case, replicate=3)
    assert not parity.compare(ref, got, exact=True)
    for a, e in ref_err.items():
        assert np.array_equal(got_err[a], np.repeat(e[:, None], 3, axis=1)), a


@pytest.mark.parametrize("name", ["p001_test10", "cbp_fullchain", "cbp_delt15", "i001_test10", "p001_midyear", "cbp_pqual",
                                  "i_iqual_variants", "p_metric_pack"])
def test_chunked_equals_single_pass(name):
    """State (named + hidden) carried through the per-segment state record makes chunked runs bit-identical."""
    case = ALL[name]
    a, ea, _ = parity.run_fused(case)
    b, eb, _ = parity.run_fused(case, chunk=97)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    for k in ea:
        assert np.array_equal(ea[k], eb[k])


@pytest.mark.parametrize("name", ["p001_test10", "cbp_fullchain", "i001_test10", "i_iqual_variants"])
def test_subchunk_tasks_equal_whole_chunk(name):
    """More tiles than resident warps (the test double holds one): a launch is cut into (tile, sub-chunk) tasks and a
    tile's state passes from task to task through the segment store -- bit-identical to one task per tile, including
    a ragged last sub-chunk and a padded last tile (70 segments = 3 tiles)."""
    case = ALL[name]
    a, ea, _ = parity.run_fused(case, replicate=70, chunk=500, sub_steps=0)
    b, eb, _ = parity.run_fused(case, replicate=70, chunk=500, sub_steps=37)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
        assert np.array_equal(a[k], np.repeat(a[k][:, :1], 70, axis=1), equal_nan=True), k
    for k in ea:
        assert np.array_equal(ea[k], eb[k])
    ref, ref_err = parity.run_oracle(case)
    assert not parity.compare(ref, b, exact=True, seg=69)


@pytest.mark.parametrize("name", ["p001_test10", "cbp_fullchain", "i001_test10", "p_stress_wet", "p_metric_pack", "i001_metric"])
def test_float32_outputs_are_the_reference_cast(name):
    """HSP2G_F32: the kernel's store rounds to float32 exactly like save_timeseries' astype(np.float32)
    (utilities.py:313) -- bit for bit on the CPU double, where the fp64 values are the oracle's."""
    if name not in ALL:
        pytest.skip("case not defined")
    case = ALL[name]
    ref, _ = parity.run_oracle(case)
    got, _, _ = parity.run_fused(case, replicate=2, out_dtype=np.float32, chunk=1000)
    for k, g in got.items():
        assert g.dtype == np.float32
        assert parity.f32_ulp_diff(g[:, 1], ref[k][:len(g)]) == 0, k


@pytest.mark.parametrize("name", sorted(ALL))
def test_dropin_bit_exact(name):
    case = ALL[name]
    ref, ref_err = parity.run_oracle(case)
    got, got_err = parity.run_dropin(case)
    assert set(ref) == set(got), "the drop-in must leave the same set of ts names as the reference wrappers"
    assert not parity.compare(ref, got, exact=True)
    for a, e in ref_err.items():
        assert np.array_equal(got_err[a], e), a


def test_iwtgas_without_its_states_raises_like_the_reference():
    """test10.uci switches IWTGAS on for IMPLND 1 but has no IWT-INIT table, and readUCI supplies no default for it: the
    reference's _iwtgas_ then raises KeyError('SOTMP') (IWTGAS.py:79; checked against the reference in the build
    container).  The drop-in refuses the same way instead of inventing a value."""
    from hspsquared_b200 import activities, test10

    case = ALL["i001_test10"]
    si = CASES.siminfo_for(case)
    si["time_unit"] = "us"
    ts = dict(CASES.forcings(case))
    ts["SURO"] = np.zeros(case["steps"])
    uci = test10.implnd()["IWTGAS"]
    assert "STATES" not in uci
    with pytest.raises(KeyError):
        activities.iwtgas(None, si, uci, ts)


@pytest.mark.parametrize("order", [None, "flags"])
def test_constituents_fed_on_the_device(order):
    """Land batch -> device gather -> PQUAL rows without a host round trip == the host-side chain, bit for bit."""
    parity.check_qual_chain_on_device(order=order)


@pytest.mark.parametrize("layout,order,jit", [("tiled", None, False), ("plain", "flags", False), ("tiled", "flags", True),
                                               ("plain", None, True)])
def test_constituents_read_the_land_buffers_in_place(layout, order, jit):
    """hsp2g_run_device_linked: the PQUAL rows, constituent-major over the land batch's tiles, read its output and forcing buffers
    in place (no gather, no scratch) == the host-side chain, bit for bit; general kernels and the ones built for the batch."""
    parity.check_qual_chain_on_device(order=order, chain="linked", layout=layout, jit=jit)


def test_abi_error_behaviour_and_smallest_shapes():
    parity.check_abi_errors_and_edges()


def test_option_word_kernels_equal_the_general_ones():
    parity.check_option_word_kernels()


def test_dispatcher_calls_are_served_from_fused_batches():
    parity.check_lazy_dispatch(n=12)


def test_dispatcher_calls_are_served_from_fused_batches_in_a_metric_run():
    """siminfo['units'] = 2 through the dispatcher: the sections behind the water budget find the converted series of the
    earlier ones in ts, exactly as the reference's wrappers leave them."""
    parity.check_lazy_dispatch(n=6, pairs=(("PERLND", "cbp_fullchain_metric"), ("IMPLND", "i_chain_metric")))


def test_metric_units_entry_conversions_of_the_chain_sections():
    """siminfo['units'] = 2 in the sections behind the water budget: what the reference converts on entry of each core is
    converted when the store is built, in its expression order (SEDMNT.py:117, PWTGAS.py:94-95, PSTEMP.py:80-101: a metric run's
    soil temperatures are deg C as they are, 16 deg C when the table is absent); the per-interval conversions are the kernels'
    (the *_metric cases of test_fused_bit_exact / test_dropin_bit_exact)."""
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.segstore import SegmentStore
    from oracle import cases as CASES

    case = CASES.build_cases()["cbp_fullchain"]
    for grp in ("PARAMETERS", "STATES"):  # no initial soil temperatures: the unit system's defaults
        for k in ("AIRTC", "SLTMP", "ULTMP", "LGTMP"):
            case["tables"]["PSTEMP"].get(grp, {}).pop(k, None)
    cal = Calendar("1976-01-01", "1976-01-03", 60, "us")
    en = SegmentStore.from_tables("PERLND", [case["tables"]], cal, units=1)
    me = SegmentStore.from_tables("PERLND", [case["tables"]], cal, units=2)
    P, S = en.P, en.S
    from hspsquared_b200.segstore import make_ui

    dets = make_ui(case["tables"]["SEDMNT"])["DETS"]
    assert me.state[S["DETS"], 0] == dets * 1.10231 / 2.471 and en.state[S["DETS"], 0] == dets
    elev = make_ui(case["tables"]["PWTGAS"])["ELEV"]
    assert me.par[P["ELEVGC"], 0] == math.pow((288.0 - 0.00198 * (elev * 3.281)) / 288.0, 5.256)
    assert en.par[P["ELEVGC"], 0] == math.pow((288.0 - 0.00198 * elev) / 288.0, 5.256)
    assert me.state[S["SLTMP"], 0] == 16.0 and en.state[S["SLTMP"], 0] == (60.0 - 32.0) * 0.555


@pytest.mark.parametrize("layout", ["plain", "tiled"])
def test_groups_by_option_word_equal_the_mixed_batch(layout):
    parity.check_grouped_batch(n=200, layout=layout)


def test_run_keeps_host_registrations_for_the_life_of_the_array():
    """LandBatch.run() page-locks ordinary numpy arrays in place; the registration is kept while the array lives (a second call
    with the same forcing / out= arrays registers nothing) and undone when the array is collected (engine._PageLocked)."""
    import gc

    from hspsquared_b200 import engine, synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    cal = Calendar("1976-01-01", "1976-01-20", 60, "us")
    ens = synth.Ensemble("PERLND", 64, seed=5, sections=("SNOW", "PWATER"))
    names = list(synth.PERLND_FORCINGS)
    f = synth.forcing_chunk(ens, cal, 0, cal.steps, names=names)
    assert f.nbytes >= engine._PageLocked.MIN_BYTES
    with LandBatch(ens.store(cal), cal, names, "all", device=0, order=None) as b:
        lib, calls = b.lib, []
        reg, unreg = lib.hsp2g_host_register, lib.hsp2g_host_unregister
        lib.hsp2g_host_register = lambda p, n: (calls.append(("reg", p)), reg(p, n))[1]
        lib.hsp2g_host_unregister = lambda p: (calls.append(("unreg", p)), unreg(p))[1]
        try:
            out = b.run(f, chunk=200)
            first = out.copy()
            assert [c for c in calls if c[0] == "reg"] == [("reg", f.ctypes.data), ("reg", out.ctypes.data)] and len(calls) == 2
            b.set_state(ens.store(cal).state)
            b.run(f, chunk=200, out=out)  # same arrays: nothing is registered again
            assert len(calls) == 2 and np.array_equal(out, first, equal_nan=True)
            fptr, optr = f.ctypes.data, out.ctypes.data
            del f, out
            gc.collect()
            assert sorted(calls[2:]) == sorted([("unreg", fptr), ("unreg", optr)])
        finally:
            lib.hsp2g_host_register, lib.hsp2g_host_unregister = reg, unreg  # the bound functions, with their argtypes
