#!/usr/bin/env python
"""The north-star run in full on one GPU: N segments (default 1,000,000) x the whole 10-year hourly run (87,672 intervals),
PERLND SNOW+PWATER, default SAVE.  The run does not fit in HBM at once (1 M segments x 240 B = 240 MB per interval), so it
is time-chunked exactly as SURVEY.md §8(d) says: each chunk's forcings are synthesised on the device (untimed, torch's
stream), the chunk's kernel launch is timed with CUDA events on the launching stream, and the outputs land in one of two
slots that are overwritten.  Reports the whole-run rate, the HBM fraction, a per-simulated-month profile of the rate, and
a bit-level check of a few segments' complete 10-year series against the CPU oracle.  Not a driver contract.

    python tools/bench_full_run.py [--segments 1000000] [--chunk 128] [--years 10] [--check 4]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=128)
    ap.add_argument("--years", type=float, default=10.0)
    ap.add_argument("--check", type=int, default=4, help="segments whose whole series is compared with the oracle")
    ap.add_argument("--caller-order", action="store_true", help="keep the generator's segment order (default: the engine's "
                    "automatic divergence-aware order, LandBatch.climate_order)")
    ap.add_argument("--layout", default="plain", choices=("plain", "tiled"))
    args = ap.parse_args()

    import torch

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from hspsquared_b200 import _lib, synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    _lib.load()
    # the calendar ends where the run ends: the reference's monthly-table interpolation depends on the stop date
    stop = bench.RUN_STOP if args.years >= 10 else np.datetime64(bench.RUN_START) + np.timedelta64(int(args.years * 8760), "h")
    cal = Calendar(bench.RUN_START, stop, 60, "ns")
    n, Tc = args.segments, args.chunk
    steps = cal.steps
    ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
    store = ens.store(cal)
    if not args.caller_order:
        sample = synth.forcing_chunk(ens, cal, 0, 24 * 7)
        perm = LandBatch.climate_order(store, list(synth.PERLND_FORCINGS), sample)
        del sample
        ens = ens.take(perm)
        store = ens.store(cal)
    save = bench.default_save()
    plain = args.layout == "plain"
    batch = LandBatch(store, cal, list(synth.PERLND_FORCINGS), save, device=0, layout=args.layout, jit=True, order=None)
    npad, nf, ns = batch.npad, batch.nf, batch.ns
    outs = bench.out_buffers(torch, npad, Tc, ns, dev, args.layout)
    stream = torch.cuda.Stream(dev)
    check = sorted(set(int(x) for x in np.linspace(0, n - 1, args.check))) if args.check else []
    got = {i: np.empty((steps, ns)) for i in check}

    chunks = [(t0, min(Tc, steps - t0)) for t0 in range(0, steps, Tc)]
    ev = []
    l0, _ = batch.kernel_stats()
    for c, (t0, m) in enumerate(chunks):
        f = bench.device_forcing(torch, ens, cal, t0, m, npad, dev, batch.forcing_rows, args.layout)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        batch.run_device(t0, m, f.data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        ev.append(e0.elapsed_time(e1))
        if plain:
            for i in check:
                got[i][t0:t0 + m] = outs[c % 2][:m, :, i].cpu().numpy()
        else:
            o = outs[c % 2].view(-1)[:(npad // 32) * m * ns * 32].view(npad // 32, m, ns, 32)  # a ragged last chunk is [tile][m]...
            for i in check:
                got[i][t0:t0 + m] = o[i // 32, :m, :, i % 32].cpu().numpy()
        del f
    l1, _ = batch.kernel_stats()
    ms = float(np.sum(ev))
    seg_steps = n * steps
    peak, peak_src = bench.measured_peak()
    achieved = bench.BYTES_PER_SEGSTEP * seg_steps / (ms * 1e-3) / 1e9

    # rate by simulated calendar month (all years pooled): where the year's average comes from
    t_mid = np.array([t0 + m // 2 for t0, m in chunks])
    stamps = (np.datetime64(bench.RUN_START) + t_mid.astype("timedelta64[h]")).astype("datetime64[M]")
    mon = stamps.astype(int) % 12
    by_month = {}
    for k in range(12):
        sel = mon == k
        if sel.any():
            w = sum(chunks[j][1] for j in np.nonzero(sel)[0]) * n
            t = float(np.sum(np.array(ev)[sel])) * 1e-3
            by_month["%02d" % (k + 1)] = {"seg_steps_per_s": w / t, "frac": bench.BYTES_PER_SEGSTEP * w / t / 1e9 / peak}

    parity = None
    if check:
        from oracle import cases as CASES
        from oracle import wrappers as Wr

        acts = {"SNOW": Wr.snow, "PWATER": Wr.pwater}
        nbad, worst, first = 0, 0.0, []
        for i in check:
            si = {"start": cal.start, "stop": cal.stop, "delt": 60, "units": 1, "steps": steps, "time_unit": "ns"}
            ts = synth.forcing_of(ens, cal, i, 0, steps)
            CASES.run_segment("PERLND", ens.tables_of(i), ts, si, acts)
            for r, full in enumerate(batch.save_rows):
                ref = ts[full.split("_", 1)[1]]
                g = got[i][:, r]
                same = (g == ref) | (np.isnan(g) & np.isnan(ref))
                nbad += int((~same).sum())
                if not same.all() and len(first) < 64:
                    k = int(np.nonzero(~same)[0][0])
                    first.append([i, full, int((~same).sum()), k, float(g[k]).hex(), float(ref[k]).hex()])
                ok = ~np.isnan(ref) & ~np.isnan(g)
                if ok.any() and not same.all():
                    den = np.maximum(np.abs(ref[ok]), 1e-4 * np.abs(ref[ok]).max() + 1e-300)
                    worst = max(worst, float((np.abs(g[ok] - ref[ok]) / den).max()))
        parity = {"segments": check, "intervals": steps, "series": ns, "values_compared": len(check) * steps * ns,
                  "values_differing": nbad, "bit_identical": nbad == 0, "max_rel_err": worst,
                  "first_differences": first}
    errs = batch.errors()
    line = {"workload": "PERLND SNOW+PWATER, %d segments x %d hourly intervals (%.2f years), default SAVE, time-chunked by %d"
                        % (n, steps, steps / 8766.0, Tc),
            "segment_order": "caller's" if args.caller_order else "automatic (LandBatch.climate_order)", "series_layout": args.layout,
            "value": seg_steps / (ms * 1e-3), "unit": "segment-timesteps/s", "kernel_seconds": ms * 1e-3,
            "launches": int(l1 - l0), "kernel": batch.kernel_name(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
            "by_month": by_month, "parity": parity,
            "error_counts_total": {k: int(np.asarray(v).sum()) for k, v in errs.items()} if isinstance(errs, dict) else None}
    batch.close()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
