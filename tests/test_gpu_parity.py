"""`-m gpu`: the CUDA library (libhsp2g.so, sm_100a) through its C ABI against the CPU oracle on the same seeded
inputs.  Bar (BASELINE.json north_star): max relative error <= 1e-9 on every saved fp64 series (NaN == NaN, -1e30
sentinels equal), error-count vectors exactly equal.  The only arithmetic difference from the oracle is CUDA's libm
(pow <= 2 ulp, exp/exp2/log <= 1 ulp) -- the device code is compiled with -fmad=false."""
import os

import numpy as np
import pytest

from oracle import cases as CASES

from . import parity

pytestmark = pytest.mark.gpu
ALL = CASES.build_cases()


@pytest.fixture(scope="module", autouse=True)
def cuda_library(oracle_lib):
    import torch

    from hspsquared_b200 import _lib

    assert torch.cuda.is_available(), "the -m gpu suite needs a CUDA device"
    lib = _lib.load()
    assert _lib.LIB_PATH.endswith(os.path.join("hspsquared_b200", "libhsp2g.so"))
    assert _lib.device_count() >= 1
    return lib


def test_device_math_equals_libm_bit_for_bit():
    """pow / exp / exp2 / log as the kernels evaluate them (glibc_math.cuh) against this box's glibc on a million
    arguments per function: every bit equal, including the special values that fall back to CUDA's libm."""
    import ctypes as C

    from hspsquared_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(2024)
    n = 1_000_000
    dp = C.POINTER(C.c_double)

    def probe(fn, x, y=None):
        out = np.empty_like(x)
        _lib.check(lib.hsp2g_math_probe(0, fn, len(x), x.ctypes.data_as(dp), y.ctypes.data_as(dp) if y is not None else None,
                                        out.ctypes.data_as(dp)))
        return out

    x = np.concatenate([rng.random(n // 2) * 3.0, np.exp((rng.random(n // 2) - 0.5) * 60.0)])
    y = np.concatenate([0.1 + rng.random(n // 2) * 4.0, (rng.random(n // 2) - 0.5) * 16.0])
    x[:8] = [0.0, 1.0, 2.0, 1e-320, np.inf, 10.0, 0.5, 1.0000001]
    y[:8] = [0.6, 5.0, 0.0, 0.6, 0.5, -400.0, 2000.0, 1.667]
    import math

    ref = np.array([math.pow(a, b) for a, b in zip(x, y)])
    got = probe(0, x, y)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), int((got.view(np.uint64) != ref.view(np.uint64)).sum())
    xe = (rng.random(n) - 0.5) * np.where(np.arange(n) % 3 == 0, 300.0, 8.0)
    assert np.array_equal(probe(1, xe).view(np.uint64), np.array([math.exp(v) for v in xe]).view(np.uint64))
    assert np.array_equal(probe(2, xe).view(np.uint64), np.array([math.exp2(v) for v in xe]).view(np.uint64))
    xl = np.where(np.arange(n) % 4 == 0, 0.9 + rng.random(n) * 0.2, np.exp((rng.random(n) - 0.5) * 200.0))
    assert np.array_equal(probe(3, xl).view(np.uint64), np.array([math.log(v) for v in xl]).view(np.uint64))


@pytest.mark.parametrize("name", sorted(ALL))
def test_fused_bit_exact_on_gpu(name):
    """With the libm restatement the CUDA path reproduces the oracle (hence the reference) BIT FOR BIT: every saved
    series of every case, error counts exactly equal.  (On a host whose libm is not the FMA build the bar is RTOL.)"""
    case = ALL[name]
    ref, ref_err = parity.run_oracle(case)
    got, got_err, kname = parity.run_fused(case, replicate=2)
    bad = parity.compare(ref, got, exact=parity.GPU_EXACT)
    assert not bad, "%s via %s: %s" % (name, kname, bad[:8])
    for a, e in ref_err.items():
        assert np.array_equal(got_err[a], np.repeat(e[:, None], 2, axis=1)), (a, got_err[a][:, 0], e)


@pytest.mark.parametrize("name", ["p001_test10", "cbp_fullchain", "cbp_delt15", "i001_test10", "p001_midyear", "cbp_pqual",
                                  "i_iqual_variants"])
def test_chunked_equals_single_pass_bitwise(name):
    case = ALL[name]
    a, ea, _ = parity.run_fused(case)
    b, eb, _ = parity.run_fused(case, chunk=97)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    for k in ea:
        assert np.array_equal(ea[k], eb[k])


@pytest.mark.parametrize("name", ["p001_test10", "i001_test10", "cbp_fullchain", "p_nosnow_iffc2", "flowtest_pwater",
                                  "flowtest_iwater", "i_atemp_snow", "i001_iwtgas", "i_solids_m1_wtf_monthly",
                                  "i_solids_iwtgas_alone", "i001_iqual", "i_iqual_variants", "cbp_pqual", "p_pqual_alone",
                                  "p_metric_pack", "i001_metric", "flowtest_pwater_metric", "cbp_fullchain_metric",
                                  "i_chain_metric", "i_solids_iwtgas_alone_metric"])
def test_dropin_vs_oracle(name):
    case = ALL[name]
    ref, ref_err = parity.run_oracle(case)
    got, got_err = parity.run_dropin(case)
    assert set(ref) == set(got)
    bad = parity.compare(ref, got, exact=parity.GPU_EXACT)
    assert not bad, bad[:8]
    for a, e in ref_err.items():
        assert np.array_equal(got_err[a], e), a


def test_golden_reference_arrays(golden_dir):
    """Directly against the committed fp64 outputs of the reference's Numba kernels (tests/golden/ref_*.npz)."""
    for name in ("p001_test10", "i001_test10", "cbp_fullchain", "p001_snow_pwater"):
        ref = np.load(os.path.join(golden_dir, "ref_%s.npz" % name))
        got, _, _ = parity.run_fused(ALL[name])
        for k, g in got.items():
            if k in ref.files:
                if parity.GPU_EXACT:  # the reference's own fp64 arrays, bit for bit
                    assert np.array_equal(g[:, 0], ref[k], equal_nan=True), (name, k, parity.rel_err(g[:, 0], ref[k]))
                else:
                    assert parity.rel_err(g[:, 0], ref[k]) <= parity.RTOL, (name, k)


def test_many_segments_padding_and_blocks():
    """333 jitter-free replicas (not a multiple of 128): every column equals column 0 bit for bit."""
    case = ALL["p001_snow_pwater"]
    got, errs, kname = parity.run_fused(case, replicate=333, chunk=1000)
    assert "SNOW|PWATER" in kname
    for k, g in got.items():
        assert np.array_equal(g, np.repeat(g[:, :1], 333, axis=1), equal_nan=True), k


@pytest.mark.parametrize("name,save", [("p001_snow_pwater", ["SNOW_PACKF", "PWATER_SURO", "PWATER_AGWS", "PWATER_UZS"]),
                                       ("i001_test10", ["SNOW_PACKF", "IWATER_SURO", "IWATER_RETS"])])
def test_subchunk_tasks_bitwise(name, save):
    """More tiles (80,000 segments = 2,500) than resident warps: the ticket scheduler hands (tile, sub-chunk) tasks
    to whichever warp is free, state passes through the segment store.  Must be bit-identical to whole-chunk tasks
    and to a 3-segment batch, whatever SM ran which sub-chunk."""
    case = ALL[name]
    n, T = 80000, 300
    a, ea, _ = parity.run_fused(case, replicate=n, chunk=130, sub_steps=0, save=save, max_steps=T)
    b, eb, _ = parity.run_fused(case, replicate=n, chunk=130, sub_steps=32, save=save, max_steps=T)
    small, es, _ = parity.run_fused(case, replicate=3, save=save, max_steps=T)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
        assert np.array_equal(b[k], np.repeat(small[k][:, :1], n, axis=1), equal_nan=True), k
    for k in ea:
        assert np.array_equal(ea[k], eb[k])
        assert np.array_equal(eb[k][:, :3], es[k])


@pytest.mark.parametrize("name", ["p001_snow_pwater", "p001_test10", "i001_test10", "cbp_fullchain", "p_metric_pack"])
def test_float32_outputs(name):
    """HSP2G_F32 (what save_timeseries stores, utilities.py:313): exactly the oracle's fp64 series cast to float32."""
    case = ALL[name]
    ref, _ = parity.run_oracle(case)
    got, _, kname = parity.run_fused(case, replicate=2, out_dtype=np.float32)
    worst = 0
    for k, g in got.items():
        assert g.dtype == np.float32
        worst = max(worst, parity.f32_ulp_diff(g[:, 0], ref[k][:len(g)]))
    assert worst <= (0 if parity.GPU_EXACT else 1), (kname, worst)


def test_batch_kernels_are_built_and_selected():
    """Every batch gets a kernel built for its own configuration (hspsquared_b200/jit.py): the option word of a uniform
    ensemble is fixed at compile time whatever the word (test10's and the CBP full-chain segment's here); the float32 and
    plain-layout variants are kernels of their own; the library's general kernels give the same bits."""
    from hspsquared_b200 import synth, test10
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    cal = Calendar("1976-01-01", "1976-01-05", 60, "ns")
    ens = synth.Ensemble("PERLND", 64, seed=5, sections=("SNOW", "PWATER"))
    save = ["SNOW_" + n for n in test10.SAVE_DEFAULT["SNOW"]] + ["PWATER_" + n for n in test10.SAVE_DEFAULT["PWATER"]]
    f = synth.forcing_chunk(ens, cal, 0, cal.steps)
    word = int(ens.store(cal).flags[0])
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), save, jit=True) as b:
        o64 = b.run(f)
        assert b.jit_error is None, b.jit_error
        k = b.kernel_name()
        assert k.startswith("perlnd_kernel<SNOW|PWATER,hourly,forc=6:") and "save=24:" in k and "plain" in k
        assert "options=%#x" % word in k and k.endswith(",jit>")
        assert b.jit_meta["spill_bytes"] == 0 and b.jit_meta["registers"] <= 128
    cbp = synth.Ensemble("PERLND", 64, seed=5, base="cbp")  # a non-test10 word: the CBP land segment's full chain
    fc = synth.forcing_chunk(cbp, cal, 0, cal.steps, names=list(synth.FULLCHAIN_FORCINGS))
    with LandBatch(cbp.store(cal), cal, list(synth.FULLCHAIN_FORCINGS), list(parity.cbp10()), jit=True) as b:
        oc = b.run(fc)
        assert b.jit_error is None, b.jit_error
        assert "options=%#x" % int(cbp.store(cal).flags[0]) in b.kernel_name() and "save=10:" in b.kernel_name()
    with LandBatch(cbp.store(cal), cal, list(synth.FULLCHAIN_FORCINGS), list(parity.cbp10()), jit=False) as b:
        assert np.array_equal(b.run(fc), oc, equal_nan=True)
        assert "jit" not in b.kernel_name()
    st = ens.store(cal)
    st.flags[-1] ^= np.uint64(1) << np.uint64(st.FB["IFRDFG"])  # one segment differs: run-time option words, same bits elsewhere
    with LandBatch(st, cal, list(synth.PERLND_FORCINGS), save, jit=True, order=None) as b:
        omix = b.run(f)
        assert "(1 bits per segment)" in b.kernel_name()  # the one bit that differs is read per segment, the rest stay fixed
    assert np.array_equal(omix[:, :, :-1], o64[:, :, :-1], equal_nan=True)
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), save, jit=True) as b:
        b.set_output_dtype(np.float32)
        b.set_forcing_dtype(np.float32)
        f32 = f.astype(np.float32)
        o32 = b.run(f32)
        assert "forc=6:0x7e,f32" in b.kernel_name() and ",f32,plain" in b.kernel_name()
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), save, jit=False, layout="tiled") as b:
        o32g = b.run(f32.astype(np.float64)).astype(np.float32)  # general kernel, tiled, widened inputs
        assert b.kernel_name() == "perlnd_kernel<SNOW|PWATER,hourly,save=default24>"
    assert np.array_equal(o32, o32g, equal_nan=True)
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), list(reversed(save)), jit=True) as b:
        orev = b.run(f)  # caller's row order is honoured whatever the device order
    assert np.array_equal(orev[:, ::-1, :], o64, equal_nan=True)


def test_categorical_fullchain_ensemble():
    """BASELINE.json configs[3] shape (full PERLND chain, option flags drawn per segment => divergent warps) on the GPU:
    sampled segments against the oracle on their own tables (<= 1e-9, error counts exact), and divergence-aware ordering
    (sort_by_flags) must only permute columns, bit for bit."""
    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    cal = Calendar("1976-02-20", "1976-03-20", 60, "us")
    steps, n = cal.steps, 5000
    ens = synth.Ensemble("PERLND", n, seed=99, base="cbp", options=("RTOPFG", "UZFG", "SNOPFG", "ICEFG", "TSOPFG"))
    names = list(synth.FULLCHAIN_FORCINGS)
    f = synth.forcing_chunk(ens, cal, 0, steps, names=names)
    with LandBatch(ens.store(cal), cal, names, "all") as b:
        out = b.run(f, chunk=250)
        save, errs = b.save, b.errors()
    assert len(save) == 84
    for i in (0, 1234, 2500, 3777, n - 1):
        si = {"start": cal.start, "stop": cal.stop, "delt": 60, "units": 1, "steps": steps, "time_unit": "us"}
        ts = synth.forcing_of(ens, cal, i, 0, steps)
        e = CASES.run_segment("PERLND", ens.tables_of(i), ts, si, parity.ORACLE_ACTS)
        for r, full in enumerate(save):
            refv = ts[full.split("_", 1)[1]]
            if parity.GPU_EXACT:
                assert np.array_equal(out[:, r, i], refv, equal_nan=True), (i, full)
            else:
                assert parity.rel_err(out[:, r, i], refv) <= parity.RTOL, (i, full)
        assert np.array_equal(e["PWATER"], np.array([errs["PWATER%d" % j][i] for j in range(10)]))
        assert np.array_equal(e["PSTEMP"], np.array([errs["PSTEMP%d" % j][i] for j in range(6)]))
    store = ens.store(cal)
    perm = store.sort_by_flags()
    with LandBatch(store, cal, names, "all") as b:
        out2 = b.run(np.ascontiguousarray(f[:, :, perm]), chunk=250)
    assert np.array_equal(out2, out[:, :, perm], equal_nan=True)


def test_water_balance_at_full_size():
    """Size-independent property at BASELINE.json configs[1] size (100,000 jittered segments): over any window the
    PWATER budget closes, SUPY + lateral inflow (0) = PERO + IGWI + TAET + change of PERS, per segment."""
    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    cal = Calendar("1976-01-01", "1976-01-09", 60, "ns")
    steps, n = cal.steps, 100000
    ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
    save = ["PWATER_SUPY", "PWATER_PERO", "PWATER_IGWI", "PWATER_TAET", "PWATER_PERS"]
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), save) as b:
        out = b.run(synth.forcing_chunk(ens, cal, 0, steps), chunk=96)
        assert not any(v.any() for v in b.errors().values())
    supy, pero, igwi, taet, pers = (out[:, r, :] for r in range(5))
    lhs = supy[1:].sum(axis=0)
    rhs = pero[1:].sum(axis=0) + igwi[1:].sum(axis=0) + taet[1:].sum(axis=0) + (pers[-1] - pers[0])
    scale = np.maximum(pers[0], 1.0)
    assert np.abs(lhs - rhs).max() / scale.max() < 1e-12, np.abs(lhs - rhs).max()


def test_run_page_locks_ordinary_numpy_arrays():
    """LandBatch.run() registers ordinary numpy arrays in place for the call (cudaHostRegister) and leaves arrays that already
    are page-locked alone; results do not depend on where the buffers live, `out=` is filled in place, and a refused
    registration leaves no error behind for the next launch."""
    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch, pinned_empty

    cal = Calendar("1976-02-01", "1976-02-15", 60, "us")
    ens = synth.Ensemble("PERLND", 3000, seed=5, sections=("SNOW", "PWATER"))
    f = synth.forcing_chunk(ens, cal, 0, cal.steps)
    assert f.nbytes > (1 << 20)
    save = ["SNOW_PACKF", "PWATER_SURO", "PWATER_AGWO", "PWATER_UZS"]
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), save, order=None) as b:
        first = b.run(f, chunk=100)
        state = b.get_state()
        fp = pinned_empty(f.shape)
        fp[...] = f
        op = pinned_empty(first.shape)
        for fin, out in ((fp, None), (f, np.empty_like(first)), (fp, op), (f[:, :, :], None)):
            b.set_state(ens.store(cal).state)
            got = b.run(fin, chunk=100, out=out)
            assert out is None or got is out
            assert np.array_equal(got, first, equal_nan=True)
            assert np.array_equal(b.get_state(), state, equal_nan=True)
        lib = b.lib
        assert lib.hsp2g_host_register(fp.ctypes.data, fp.nbytes) != 0 and b"cudaHostRegister" in lib.hsp2g_last_error()
        b.set_state(ens.store(cal).state)
        assert np.array_equal(b.run(f, chunk=100), first, equal_nan=True)  # no stale CUDA error after the refusal


def test_shared_met_mode_on_gpu():
    """Shared-met mode (cp.async gathers from the station table + MFACTOR at the point of use) == per-segment forcing
    arrays through the TMA ring, bit for bit, on 2,000 segments."""
    parity.check_shared_met_mode(n=2000)


def test_shared_met_batch_kernels_table_and_gather_on_gpu():
    """The kernels built for the batch in both shared-met modes: per-lane cp.async gathers (the default) and, with
    HSP2G_MET_TABLE=1 and <= 32 stations, one bulk copy of the interval's station table per warp == per-segment forcing
    arrays, bit for bit."""
    parity.check_shared_met_mode(n=2500, jit=True)
    parity.check_shared_met_mode(n=2500, jit=True, nst=37)
    parity.check_shared_met_mode(n=2500, jit=True, table=True)
    parity.check_shared_met_mode(n=2500, jit=None, table=True)


def test_metric_units_compiled_into_batch_kernels_on_gpu():
    parity.check_metric_batch_kernels(replicate=70)


def test_hspf_known_answers(golden_dir):
    """The reference's own KATs through the GPU: Flow_Test.plt (test_ipwater.py tolerance) and test10I.hbn."""
    hspf = np.load(os.path.join(golden_dir, "hspf_flowtest.npz"))
    got, errs, _ = parity.run_fused(ALL["flowtest_pwater"])
    assert not errs["PWATER"].any()
    for v in ("SURO", "IFWO", "AGWO"):
        assert abs((got[v][:, 0] - hspf["PERLND1_" + v]).sum()) < 1e-3
    got, errs, _ = parity.run_fused(ALL["flowtest_iwater"])
    assert not errs["IWATER"].any()
    assert abs((got["SURO"][:, 0] - hspf["IMPLND1_SURO"]).sum()) < 1e-3
    h10 = np.load(os.path.join(golden_dir, "hspf_test10I.npz"))
    got, _, _ = parity.run_fused(ALL["i001_test10"])
    for v, t in {"SUPY": 2e-6, "SURO": 2e-6, "SURS": 1e-6, "RETS": 1e-6, "PET": 1e-7, "IMPEV": 1e-7}.items():
        assert np.abs(got[v][:, 0] - h10[v]).max() < t, v
    # SOLIDS and IWTGAS groups of the same HSPF run, through the fused IMPLND kernel
    got, _, kname = parity.run_fused(ALL["i001_iwtgas"], save=["SOLIDS_SLDS", "SOLIDS_SOSLD", "IWTGAS_SOTMP", "IWTGAS_SODOX",
                                                                "IWTGAS_SOCO2", "IWTGAS_SOHT", "IWTGAS_SODOXM", "IWTGAS_SOCO2M"])
    for key in ("SOLIDS_SLDS", "SOLIDS_SOSLD", "IWTGAS_SOTMP", "IWTGAS_SODOX", "IWTGAS_SOCO2", "IWTGAS_SOHT",
                "IWTGAS_SODOXM", "IWTGAS_SOCO2M"):
        ref, g = h10[key], got[key.split("_", 1)[1]][:, 0]
        ok = ref > -1e29
        assert np.array_equal(ok, g > -1e29), key
        assert np.abs(g[ok] - ref[ok]).max() <= 2e-6 * np.abs(ref[ok]).max(), key
    # IQUAL group (constituent COD) of the same HSPF run: IMPLND kernel -> qual kernel
    got, errs, _ = parity.run_fused(ALL["i001_iqual"])
    assert not errs["IQUAL"].any()
    for cons in ("SQO", "SOQSP", "SOQOC", "SOQC", "SOQS", "SOQO", "SOQUAL"):
        ref, g = h10["IQUAL_" + cons], got["IQUAL1_" + cons][:, 0]
        ok = ref > -1e29
        assert np.array_equal(ok, g > -1e29), cons
        assert np.abs(g[ok] - ref[ok]).max() <= (1e-5 if cons in ("SOQOC", "SOQC") else 2e-6) * np.abs(ref[ok]).max(), cons


def test_qual_rows_many_segments():
    """PQUAL as a batch of (segment, constituent) rows: 5,000 jittered full-chain segments x 3 constituents fed by the land
    batch's outputs; sampled segments equal the oracle's chain, and rows of different segments do not mix."""
    from hspsquared_b200 import synth, test10
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch, QualBatch
    from hspsquared_b200.qualstore import PQUAL_INPUTS

    cal = Calendar("1976-02-01", "1976-04-01", 60, "us")
    n = 5000
    ens = synth.Ensemble("PERLND", n, seed=17, base="cbp", options=("RTOPFG", "UZFG"))
    names = list(synth.FULLCHAIN_FORCINGS)
    f = synth.forcing_chunk(ens, cal, 0, cal.steps, names=names)
    need = ["PWATER_SURO", "PWATER_IFWO", "PWATER_AGWO", "PWATER_PERO", "SEDMNT_WSSD", "SEDMNT_SCRSD"]
    with LandBatch(ens.store(cal), cal, names, need) as b:
        land = b.run(f, chunk=480)
    series = {full.split("_", 1)[1]: land[:, r, :] for r, full in enumerate(need)}
    series["PREC"] = f[:, names.index("PREC"), :]
    qt = test10.pqual_tables()
    qn = [k for k in PQUAL_INPUTS if k in series]
    with QualBatch("PERLND", [qt] * n, cal, qn, "all") as q:
        assert q.n == 3 * n
        out = q.run(q.gather(series), chunk=480)
        rows, save = q.rows, q.save
        assert "qual_kernel<PQUAL" in q.kernel_name()
    for i in (0, 2500, 4999):
        t = ens.tables_of(i)
        t["ACTIVITY"]["PQUAL"] = 1
        t["PQUAL"] = test10.pqual_tables()
        si = {"start": cal.start, "stop": cal.stop, "delt": 60, "units": 1, "steps": cal.steps, "time_unit": "us"}
        ts = synth.forcing_of(ens, cal, i, 0, cal.steps)
        CASES.run_segment("PERLND", t, ts, si, parity.ORACLE_ACTS)
        for j, (seg, k) in enumerate(rows):
            if seg != i:
                continue
            for r, full in enumerate(save):
                key = "PQUAL%d_%s" % (k, full.split("_", 1)[1])
                if parity.GPU_EXACT:
                    assert np.array_equal(out[:, r, j], ts[key], equal_nan=True), (i, key)
                else:
                    assert parity.rel_err(out[:, r, j], ts[key]) <= parity.RTOL, (i, key)


def test_period_aggregation_equals_pandas():
    """outstep 3 / 4 / 5 statistics on the device == pandas on the float32 frames (IOManager.write_ts, io.py:74-94)."""
    parity.check_period_aggregation(n=300, days=400)


def test_option_word_kernels_equal_the_general_ones():
    parity.check_option_word_kernels()


def test_abi_error_behaviour_and_smallest_shapes():
    parity.check_abi_errors_and_edges()


def test_constituents_fed_on_the_device():
    """Land kernel -> gather kernel -> qual kernel with every buffer in HBM: same bits as the host-side chain (3,000 segments,
    divergence-ordered land batch)."""
    parity.check_qual_chain_on_device(n=3000, chunk=256, order="flags")


@pytest.mark.parametrize("layout,jit", [("tiled", True), ("plain", True), ("plain", False)])
def test_constituents_read_the_land_buffers_in_place(layout, jit):
    """hsp2g_run_device_linked on the GPU: the PQUAL rows fetch their seven input rows with one 256-byte bulk copy each straight
    from the land batch's output / forcing buffers (3,000 segments -> 9,024 rows incl. the padding lanes)."""
    parity.check_qual_chain_on_device(n=3000, chunk=256, order="flags", chain="linked", layout=layout, jit=jit)


def test_metstore_feeds_shared_met_kernels():
    """EXT SOURCES rows -> MetStore (hspsquared_b200/ingest.py) -> hsp2g_set_shared_forcing: bit-identical to per-segment arrays."""
    parity.check_metstore_shared()


def test_watershed_wdm_to_frames(tmp_path):
    """BASELINE.json configs[4] in miniature on the GPU: 3,000 full-chain segments on 12 gages of a WDM file, three simulated
    months, float32 frames out; sampled segments equal the oracle."""
    parity.check_watershed(tmp_path, n=3000, gages=12, days=90, sample=(0, 1234, 2999), chunk=720)


def test_doe_on_gpu():
    """mainDoE's parameter sweep (mainDoE.py:19-140) as one batch of (run, segment) rows through the CUDA library: every run of
    the reference's own docstring example equals the oracle on that run's tables."""
    from .test_doe import DOE_LIST

    parity.check_doe(DOE_LIST)


def test_dispatcher_calls_are_served_from_fused_batches():
    """SURVEY 8b "granularity mismatch" on the GPU: (segment, activity) calls of a dispatcher shaped like hsp2.main are served
    from one fused batch per operation (hspsquared_b200/lazybatch.py), every series and error vector equal to the oracle's."""
    parity.check_lazy_dispatch(n=40)


def test_dispatcher_calls_are_served_from_fused_batches_in_a_metric_run():
    """... and with siminfo['units'] = 2: the fused chain hands SEDMNT / PSTEMP / PWTGAS / SOLIDS / IWTGAS the converted series."""
    parity.check_lazy_dispatch(n=20, pairs=(("PERLND", "cbp_fullchain_metric"), ("IMPLND", "i_chain_metric")))


def test_results_stream_into_the_product_store(tmp_path):
    """north_star subsystem 4 on the GPU: SAVEd series leave the device chunk by chunk into pinned buffers and a writer thread
    puts them into the NpyResults store (a SupportsWriteTS / SupportsReadTS backend) while the next chunk computes; every frame
    read back equals what save_timeseries would have written (2,500 segments in a device order that differs from the caller's)."""
    parity.check_streaming_sink(tmp_path, n=2500, days=20, chunk=100, pinned=True)


def test_handles_driven_from_threads_of_one_process():
    """SURVEY 8b "Threading": one handle per thread, several threads of ONE process (the reference's kernels hold the GIL; this
    library's calls release it in ctypes).  Six threads each own a batch on cuda:0 and run it concurrently -- creation, uploads,
    kernel builds, chunked runs, state and error read-back all overlap -- and every thread's result equals the same batch run
    alone."""
    import threading

    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    cal = Calendar("1976-02-01", "1976-02-11", 60, "us")
    jobs = []
    for k in range(6):
        op, secs = (("PERLND", ("SNOW", "PWATER")), ("IMPLND", ("SNOW", "IWATER")))[k % 2]
        ens = synth.Ensemble(op, 3000 + 37 * k, seed=40 + k, sections=secs)
        jobs.append((ens, synth.forcing_chunk(ens, cal, 0, cal.steps)))

    def run(job, jit):
        ens, f = job
        with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), "all", jit=jit) as b:
            out = b.run(f, chunk=60)
            return out, b.get_state(), b.errors()

    alone = [run(j, False) for j in jobs]
    got, errs = [None] * len(jobs), []

    def worker(i):
        try:
            got[i] = run(jobs[i], True)
        except Exception as ex:  # pragma: no cover
            errs.append((i, repr(ex)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errs, errs
    for (o, st, er), (o2, st2, er2) in zip(alone, got):
        assert np.array_equal(o, o2, equal_nan=True) and np.array_equal(st, st2, equal_nan=True)
        assert all(np.array_equal(er[k], er2[k]) for k in er)


@pytest.mark.parametrize("layout", ["plain", "tiled"])
def test_groups_by_option_word_equal_the_mixed_batch(layout):
    """BASELINE configs[3] machinery: a full-chain ensemble with per-segment option flags as one batch (and one kernel built for
    its word) per option word, launched together on the device, equals the single mixed batch bit for bit (3,000 segments)."""
    parity.check_grouped_batch(n=3000, layout=layout, jit=True)


@pytest.mark.parametrize("layout", ["plain", "tiled"])
def test_overlapped_launches_are_bit_identical(layout):
    """hsp2g_set_launch_overlap: consecutive launches of one batch overlap at their tails (programmatic dependent launch, per-tile
    hand-over through progress[tile]).  100,000 segments = 3,125 tiles > resident warps, so the grid fills the device and the
    launches really are chained; every saved value, the final state and the error counts equal the serialised run's."""
    import torch

    import bench
    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    dev = torch.device("cuda", 0)
    cal = Calendar("1976-02-10", "1976-03-05", 60, "ns")
    n, Tc, chunks = 100_000, 96, 6
    assert Tc * chunks <= cal.steps
    ens = synth.Ensemble("PERLND", n, seed=21, sections=("SNOW", "PWATER"))
    names = list(synth.PERLND_FORCINGS)
    save = ["SNOW_PACKF", "SNOW_WYIELD", "PWATER_SURO", "PWATER_AGWS", "PWATER_UZS", "PWATER_PERO"]
    stream = torch.cuda.Stream(dev)
    results = []
    for overlap in (True, False):
        with LandBatch(ens.store(cal), cal, names, save, layout=layout, order=None) as b:
            forc = [bench.device_forcing(torch, ens, cal, c * Tc, Tc, b.npad, dev, b.forcing_rows, layout) for c in range(chunks)]
            outs = bench.out_buffers(torch, b.npad, Tc, b.ns, dev, layout, count=chunks)
            b.set_schedule(32)
            b.set_launch_overlap(overlap)
            torch.cuda.synchronize(dev)
            for c in range(chunks):
                b.run_device(c * Tc, Tc, forc[c].data_ptr(), outs[c].data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize(dev)
            assert ",jit>" in b.kernel_name()
            assert b.chained_launches() == (chunks - 1 if overlap else 0)
            results.append(([o.cpu().numpy() for o in outs], b.get_state(), b.errors()))
            del forc, outs
    (oa, sa, ea), (ob, sb, eb) = results
    for a, b_ in zip(oa, ob):
        assert np.array_equal(a, b_, equal_nan=True)
    assert np.array_equal(sa, sb, equal_nan=True)
    assert all(np.array_equal(ea[k], eb[k]) for k in ea)
