#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one GPU (device-resident forcings, CUDA events on the launching
stream) -- the numbers quoted in profiles/ and DESIGN.md beside the headline that bench.py measures.  Not a driver contract.

    python tools/bench_configs.py [implnd] [fullchain] [fullchain_sorted] [perlnd_1m] [perlnd_flows] ...
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (device_forcing, measured_peak, default_save)


def run_qual(name, torch, dev):
    """PQUAL / IQUAL rows: `n` (segment, constituent) rows reading synthetic flow / sediment series resident in HBM."""
    from hspsquared_b200 import test10
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import QualBatch
    from hspsquared_b200.qualstore import IQUAL_INPUTS, PQUAL_INPUTS

    cal = Calendar(bench.RUN_START, bench.RUN_STOP, 60, "ns")
    K, W, Tc = 8, 3, 256
    imp = name == "iqual"
    nseg = 100_000
    tabs = test10.iqual_tables() if imp else test10.pqual_tables()
    names = list(IQUAL_INPUTS[:3] if imp else PQUAL_INPUTS[:7])
    with QualBatch("IMPLND" if imp else "PERLND", [tabs] * nseg, cal, names, "all", device=dev.index or 0, layout="tiled") as batch:
        npad, nf, ns, n = batch.npad, batch.nf, batch.ns, batch.n
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        forc = []
        for c in range(K + W):
            # storms are shared by all rows (7 % of the intervals), magnitudes differ per row; dry intervals carry no
            # surface runoff, sediment or solids, subsurface flows are small and always there
            u = torch.rand((npad // 32, Tc, nf, 32), generator=g, device=dev, dtype=torch.float64)
            wet = (torch.rand((1, Tc, 1), generator=g, device=dev) > 0.93).to(torch.float64)
            f = torch.empty_like(u)
            if imp:  # SURO SOSLD PREC
                f[:, :, 0] = wet * 0.03 * u[:, :, 0]
                f[:, :, 1] = wet * 0.002 * u[:, :, 1]
                f[:, :, 2] = wet * 0.05 * u[:, :, 2]
            else:  # SURO IFWO AGWO PERO WSSD SCRSD PREC
                f[:, :, 0] = wet * 0.02 * u[:, :, 0]
                f[:, :, 1] = 0.001 * u[:, :, 1]
                f[:, :, 2] = 0.001 + 0.001 * u[:, :, 2]
                f[:, :, 3] = f[:, :, 0] + f[:, :, 1] + f[:, :, 2]
                f[:, :, 4] = wet * 0.05 * u[:, :, 4]
                f[:, :, 5] = wet * 0.01 * u[:, :, 5] * (u[:, :, 3] > 0.5)
                f[:, :, 6] = wet * 0.05 * u[:, :, 6]
            forc.append(f)
            del u
        outs = [torch.empty((npad // 32, Tc, ns, 32), dtype=torch.float64, device=dev) for _ in range(2)]
        stream = torch.cuda.Stream(dev)
        torch.cuda.synchronize(dev)
        for c in range(W):
            batch.run_device(c * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for c in range(W, W + K):
            batch.run_device(c * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        peak, _src = bench.measured_peak()
        B = 8 * (nf + ns)
        value = n * Tc * K / (ms * 1e-3)
        line = {"config": name, "kernel": batch.kernel_name(), "rows": n, "segments": nseg, "intervals_per_step": Tc, "steps": K,
                "forcing_rows": nf, "saved_rows": ns, "bytes_per_row_interval": B, "ms_per_step": ms / K,
                "row_timesteps_per_s": value, "algorithmic_GBps": value * B / 1e9, "frac_of_measured_hbm": value * B / 1e9 / peak,
                "note": "one row = one (segment, constituent) pair; synthetic flow / sediment series: storms in 7 % of the intervals, "
                        "shared by all rows, per-row magnitudes"}
    del forc, outs
    torch.cuda.empty_cache()
    return line


def run_chain(name, torch, dev):
    """configs[3] with its quality constituents: 250 k full-chain segments (6 flow / sediment series saved) and 3 PQUAL constituents
    per segment (750 k rows, all 20 outputs each), everything resident in HBM; timed end to end per chunk.
    fullchain_pqual[_plain]: the constituents read the land batch's buffers in place (hsp2g_run_device_linked);
    fullchain_pqual_gather: round 1's chain, a device gather into a scratch buffer between the two kernels."""
    from hspsquared_b200 import synth, test10
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch, QualBatch
    from hspsquared_b200.qualstore import PQUAL_INPUTS

    cal = Calendar(bench.RUN_START, bench.RUN_STOP, 60, "ns")
    K, W, Tc, n = 8, 3, 128, 250_000
    gather = name.endswith("_gather")
    layout = "plain" if name.endswith("_plain") else "tiled"
    ens = synth.Ensemble("PERLND", n, seed=1234, base="cbp", options=("RTOPFG", "UZFG", "SNOPFG", "ICEFG", "TSOPFG"))
    forcings = list(synth.FULLCHAIN_FORCINGS)
    need = ["PWATER_SURO", "PWATER_IFWO", "PWATER_AGWO", "PWATER_PERO", "SEDMNT_WSSD", "SEDMNT_SCRSD"]
    qt = test10.pqual_tables()
    with LandBatch(ens.store(cal), cal, forcings, need, device=dev.index or 0, order="flags", layout=layout) as land, \
            QualBatch("PERLND", [qt] * n, cal, list(PQUAL_INPUTS[:7]), "all", device=dev.index or 0, land=land, layout=layout,
                      chain="gather" if gather else "linked") as q:
        ens2 = ens.take(land.order)  # forcings generated in the land batch's device order
        forc = [bench.device_forcing(torch, ens2, cal, c * Tc, Tc, land.npad, dev, land.forcing_rows, layout) for c in range(K + W)]
        lo = torch.empty(land.npad * Tc * land.ns, dtype=torch.float64, device=dev)
        qf = torch.empty(q.npad * Tc * q.nf if gather else 1, dtype=torch.float64, device=dev)
        qo = [torch.empty(q.npad * Tc * q.ns, dtype=torch.float64, device=dev) for _ in range(2)]
        stream = torch.cuda.Stream(dev)
        torch.cuda.synchronize(dev)

        def step(c):
            land.run_device(c * Tc, Tc, forc[c].data_ptr(), lo.data_ptr(), stream.cuda_stream)
            if gather:
                q.run_device_from(c * Tc, Tc, forc[c].data_ptr(), lo.data_ptr(), qf.data_ptr(), qo[c % 2].data_ptr(), stream.cuda_stream)
            else:
                q.run_device_linked_from(c * Tc, Tc, forc[c].data_ptr(), lo.data_ptr(), qo[c % 2].data_ptr(), stream.cuda_stream)

        for c in range(W):
            step(c)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for c in range(W, W + K):
            step(c)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        peak, _src = bench.measured_peak()
        # algorithmic bytes per segment-interval: land 6 forcings in + 6 series out, constituents 7 in + 20 out per row (+ the
        # gather's 7 in + 7 out per row where there is one)
        B = 8 * (6 + 6) + 3 * 8 * (7 + 20) + (3 * 8 * (7 + 7) if gather else 0)
        value = n * Tc * K / (ms * 1e-3)
        last = qo[(W + K - 1) % 2]
        r = q.save_rows.index("PQUAL_SOQUAL")
        sq = last.view(q.npad // 32, Tc, q.ns, 32)[:, :, r].permute(1, 0, 2).reshape(Tc, -1) if layout == "tiled" \
            else last.view(Tc, q.ns, q.npad)[:, r]
        sq = sq[:, :q.n]  # [interval][device row]
        if not gather:  # the rows of the land batch's padding lanes (copies of its last segment) are not part of the result
            real = torch.tensor([sgm >= 0 for sgm, _k in q.rows], device=dev)[torch.as_tensor(q.order, device=dev)]
            sq = sq[:, real]
        line = {"config": name, "kernels": [land.kernel_name()] + (["gather_kernel"] if gather else []) + [q.kernel_name()],
                "segments": n, "constituent_rows": int(sum(1 for sgm, _k in q.rows if sgm >= 0)), "intervals_per_step": Tc, "steps": K,
                "ms_per_step": ms / K, "segment_timesteps_per_s": value, "bytes_per_segment_interval": B,
                "algorithmic_GBps": value * B / 1e9, "frac_of_measured_hbm": value * B / 1e9 / peak,
                "checksum_soqual_last_chunk": float(sq.sum().item()),
                "note": "full chain + 3 PQUAL constituents per segment, %s launches per chunk on one stream, no host round trip"
                        % ("three" if gather else "two")}
    del forc, lo, qf, qo
    torch.cuda.empty_cache()
    return line


def run_daily_e2e(name, torch, dev):
    """Headline ensemble end to end with DAILY output (outstep 3): pinned host forcings in, float32 series on the device, period
    statistics on the device (hsp2g_aggregate_f32), and only last / sum / mean per day copied back -- 3/24 of the hourly bytes."""
    import ctypes as C

    import pandas as pd

    from hspsquared_b200 import _lib, synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch
    from hspsquared_b200.results import period_bins

    lib = _lib.load()
    shared = name == "perlnd_daily_shared_e2e"  # forcings from 16 met stations x MFACTOR: nothing to copy in
    cal = Calendar(bench.RUN_START, bench.RUN_STOP, 60, "ns")
    K, W, Tc, n = 12, 3, 432, 100_000  # 432 = 18 days: chunks end where days end (labels are interval ends: offset by one)
    ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
    save = bench.default_save()
    tindex = pd.date_range(bench.RUN_START, periods=(K + W) * Tc + 1, freq="60min")[1:]
    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), save, device=dev.index or 0, layout="tiled") as b:
        b.set_output_dtype(np.float32)
        npad, nf, ns = b.npad, b.nf, b.ns
        if shared:
            nst, total = 16, (K + W) * Tc
            s0 = synth.shared_series("PERLND", cal, 0, total)
            series = np.empty((total, nf, nst))
            for k in range(nst):
                for r, nm in enumerate(synth.PERLND_FORCINGS):
                    a = s0[nm].copy()
                    if nm == "AIRTMP":
                        a = a + (k - nst / 2) * 0.4
                    elif nm == "DTMPG":
                        a = (s0["AIRTMP"] + (k - nst / 2) * 0.4) - 5.0
                    series[:, r, k] = a
            station = np.minimum((np.argsort(np.argsort(ens.u_airt)) * nst // n), nst - 1).astype(np.int32)
            mfac = np.ones((nf, n))
            mfac[list(synth.PERLND_FORCINGS).index("PREC")] = 1.0 + 0.2 * ens.u_prec
            mfac[list(synth.PERLND_FORCINGS).index("PETINP")] = 1.0 + 0.1 * ens.u_pet
            b.set_shared_forcing(series, station, mfac)
            hf = df_ = None
        else:
            hf = [torch.empty((npad // 32, Tc, nf, 32), dtype=torch.float64).pin_memory() for _ in range(2)]
            for k in range(2):
                hf[k].copy_(bench.device_forcing(torch, ens, cal, k * Tc, Tc, npad, dev, b.forcing_rows, "tiled").cpu())
            df_ = [torch.empty((npad // 32, Tc, nf, 32), dtype=torch.float64, device=dev) for _ in range(2)]
        do = torch.empty((npad // 32, Tc, ns, 32), dtype=torch.float32, device=dev)
        bins = [period_bins(tindex[c * Tc:(c + 1) * Tc], 3)[0] for c in range(K + W)]
        nb = max(int(x[-1]) + 1 for x in bins)
        da = torch.empty((npad // 32, nb, ns, 3, 32), dtype=torch.float32, device=dev)
        ha = [torch.empty((npad // 32, nb, ns, 3, 32), dtype=torch.float32).pin_memory() for _ in range(2)]
        s_in, stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)  # copy-in runs ahead of kernel + statistics + copy-out
        i32 = C.POINTER(C.c_int32)
        ready = [torch.cuda.Event() for _ in range(2)]   # forcing slot filled
        freed = [torch.cuda.Event() for _ in range(2)]   # forcing slot consumed by its kernel
        copied = [torch.cuda.Event() for _ in range(2)]  # host result slot written

        def step(c):
            k = c % 2
            if not shared:
                with torch.cuda.stream(s_in):
                    s_in.wait_event(freed[k])
                    df_[k].copy_(hf[k], non_blocking=True)
                    ready[k].record(s_in)
            with torch.cuda.stream(stream):
                if shared:
                    b.run_device(c * Tc, Tc, 0, do.data_ptr(), stream.cuda_stream)
                else:
                    stream.wait_event(ready[k])
                    b.run_device(c * Tc, Tc, df_[k].data_ptr(), do.data_ptr(), stream.cuda_stream)
                    freed[k].record(stream)
                x = np.ascontiguousarray(bins[c])
                _lib.check(lib.hsp2g_aggregate_f32(dev.index or 0, do.data_ptr(), npad // 32, Tc, ns, x.ctypes.data_as(i32), nb,
                                                   da.data_ptr(), stream.cuda_stream))
                ha[k].copy_(da, non_blocking=True)
                copied[k].record(stream)

        for c in range(W):
            step(c)
        torch.cuda.synchronize(dev)
        import time
        t0 = time.perf_counter()
        for c in range(W, W + K):
            step(c)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        chk = float(ha[(W + K - 1) % 2][:, :, 13:, 1].sum())
        value = n * Tc * K / dt
        line = {"config": name, "kernel": b.kernel_name(), "segments": n, "intervals_per_step": Tc, "steps": K,
                "ms_per_step": 1e3 * dt / K, "segment_timesteps_per_s": value,
                "h2d_bytes_per_step": 0 if shared else Tc * nf * npad * 8, "d2h_bytes_per_step": nb * ns * 3 * npad * 4, "result_checksum": chk,
                "frac_of_measured_hbm": None,
                "note": "copy-in stream + compute stream: H2D of fp64 forcings, kernel with float32 stores, hsp2g_aggregate_f32 (daily last / "
                        "sum / mean), D2H of the statistics; wall clock around the timed steps"}
    return line


def run(name, torch, dev):
    if name in ("perlnd_daily_e2e", "perlnd_daily_shared_e2e"):
        return run_daily_e2e(name, torch, dev)
    if name in ("pqual", "iqual"):
        return run_qual(name, torch, dev)
    if name.startswith("fullchain_pqual"):
        return run_chain(name, torch, dev)
    from hspsquared_b200 import synth, test10
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    cal = Calendar(bench.RUN_START, bench.RUN_STOP, 60, "ns")
    K, W, Tc = 8, 3, 256
    note = ""
    year = name.endswith("_year")  # 20 timed steps x 438 intervals = one simulated year, like bench.py
    if year:
        name = name[:-5]
    if name == "implnd":  # IMPLND SNOW+IWATER, default SAVE (13 + 3)
        n = 100_000
        ens = synth.Ensemble("IMPLND", n, seed=1234, sections=("SNOW", "IWATER"))
        forcings = list(synth.IMPLND_FORCINGS)
        save = ["SNOW_" + x for x in test10.SAVE_DEFAULT["SNOW"]] + ["IWATER_" + x for x in test10.SAVE_DEFAULT["IWATER"]]
    elif name in ("fullchain", "fullchain_sorted", "fullchain_cbp10"):  # configs[3]
        n = 250_000
        ens = synth.Ensemble("PERLND", n, seed=1234, base="cbp", options=("RTOPFG", "UZFG", "SNOPFG", "ICEFG", "TSOPFG"))
        forcings = list(synth.FULLCHAIN_FORCINGS)
        save = "all"
        if name == "fullchain_cbp10":
            save = ["PWATER_SURO", "PWATER_IFWO", "PWATER_AGWO", "SEDMNT_SOSED", "PWTGAS_SOHT", "PWTGAS_IOHT",
                    "PWTGAS_AOHT", "PWTGAS_SODOXM", "PWTGAS_IODOXM", "PWTGAS_AODOXM"]
        Tc = 128
    elif name == "perlnd_1m":  # north_star target size: 1 M segments, SNOW+PWATER default SAVE
        n = 1_000_000
        ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
        forcings, save, Tc = list(synth.PERLND_FORCINGS), bench.default_save(), 64
    elif name == "perlnd_flows":  # CBP-style: only the three runoff components leave the device
        n = 100_000
        ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
        forcings, save = list(synth.PERLND_FORCINGS), ["PWATER_SURO", "PWATER_IFWO", "PWATER_AGWO"]
    elif name == "perlnd_shared_met":  # headline segments, forcings from 16 met stations + per-segment MFACTOR (own row: never mixed in)
        n = 100_000
        ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
        forcings, save = list(synth.PERLND_FORCINGS), bench.default_save()
    elif name == "perlnd_all":
        n = 100_000
        ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
        forcings, save = list(synth.PERLND_FORCINGS), "all"
    else:
        raise SystemExit("unknown config %r" % name)
    if year:
        K, Tc = 20, 438
        note = (note + "; " if note else "") + "one simulated year (20 x 438 intervals)"
    store = ens.store(cal)
    perm = None
    if name == "fullchain_sorted":
        perm = store.sort_by_flags()
        ens = ens.take(perm)
        note = "segments ordered by option word (SegmentStore.sort_by_flags)"
    batch = LandBatch(store, cal, forcings, save, device=dev.index or 0, layout="tiled")
    npad, nf, ns = batch.npad, batch.nf, batch.ns
    shared = name == "perlnd_shared_met"
    if shared:
        nst = 16
        total = (K + W) * Tc
        s0 = synth.shared_series("PERLND", cal, 0, total)
        series = np.empty((total, nf, nst))
        for k in range(nst):
            for r, nm in enumerate(forcings):
                a = s0[nm].copy()
                if nm == "AIRTMP":
                    a = a + (k - nst / 2) * 0.4
                elif nm == "DTMPG":
                    a = (s0["AIRTMP"] + (k - nst / 2) * 0.4) - 5.0
                series[:, r, k] = a
        station = np.minimum((np.argsort(np.argsort(ens.u_airt)) * nst // n), nst - 1).astype(np.int32)
        mfac = np.ones((nf, n))
        mfac[forcings.index("PREC")] = 1.0 + 0.2 * ens.u_prec
        mfac[forcings.index("PETINP")] = 1.0 + 0.1 * ens.u_pet
        batch.set_shared_forcing(series, station, mfac)
        note = "forcings: %d met stations x MFACTOR (hsp2g_set_shared_forcing); no per-segment forcing in HBM" % nst

        class _Null:
            def data_ptr(self):
                return 0
        forc = [_Null() for _ in range(K + W)]
    else:
        forc = [bench.device_forcing(torch, ens, cal, c * Tc, Tc, npad, dev, batch.forcing_rows, "tiled") for c in range(K + W)]
    outs = [torch.empty((npad // 32, Tc, ns, 32), dtype=torch.float64, device=dev) for _ in range(2)]
    stream = torch.cuda.Stream(dev)
    torch.cuda.synchronize(dev)
    for c in range(W):
        batch.run_device(c * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for c in range(W, W + K):
        batch.run_device(c * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    peak, _src = bench.measured_peak()
    B = 8 * ((0 if shared else nf) + ns)
    value = n * Tc * K / (ms * 1e-3)
    line = {"config": name, "kernel": batch.kernel_name(), "segments": n, "intervals_per_step": Tc, "steps": K,
            "forcing_rows": nf, "saved_rows": ns, "bytes_per_segment_interval": B, "ms_per_step": ms / K,
            "segment_timesteps_per_s": value, "algorithmic_GBps": value * B / 1e9, "frac_of_measured_hbm": value * B / 1e9 / peak,
            "errors_nonzero": {k: int((v != 0).sum()) for k, v in batch.errors().items() if (v != 0).any()}, "note": note}
    batch.close()
    del forc, outs
    torch.cuda.empty_cache()
    return line


def main():
    import torch

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    for name in sys.argv[1:] or ["implnd", "perlnd_flows", "perlnd_all", "perlnd_shared_met", "fullchain", "fullchain_sorted", "fullchain_cbp10", "perlnd_1m"]:
        print(json.dumps(run(name, torch, dev)), flush=True)


if __name__ == "__main__":
    main()
