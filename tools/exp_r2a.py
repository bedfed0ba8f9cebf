#!/usr/bin/env python
"""Round-2 experiment A (one B200): the headline configuration (100k PERLND SNOW+PWATER segments, default SAVE, one simulated
year in 20 chunks of 438 intervals) timed device-resident for every kernel / layout variant, then end to end from plain host
buffers.  Writes one JSON object per line to stdout."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    import torch

    from hspsquared_b200 import _lib, synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch, pinned_empty

    n = int(os.environ.get("EXP_SEGMENTS", "100000"))
    Tc, K, W = 438, 20, 3
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cal = Calendar(B.RUN_START, B.RUN_STOP, 60, "ns")
    ens = synth.Ensemble("PERLND", n, seed=1234, sections=("SNOW", "PWATER"))
    store = ens.store(cal)
    save = B.default_save()
    peak = B.measured_peak()[0]
    stream = torch.cuda.Stream(dev)
    rows = tuple(sorted(synth.PERLND_FORCINGS, key=lambda k: _lib.index("forcings")[k]))
    npad = (n + 31) // 32 * 32
    plain64 = [B.device_forcing(torch, ens, cal, c * Tc, Tc, npad, dev, rows, "plain") for c in range(W + K)]
    torch.cuda.synchronize()

    def to_layout(f, layout, dtype):
        if dtype is not None:
            f = f.to(dtype)
        if layout == "tiled":
            return f.view(Tc, len(rows), npad // 32, 32).permute(2, 0, 1, 3).contiguous()
        return f

    ref_out = {}

    def run(label, layout, jit, f32in=False, f32out=False, env=None, word=True, sub=None):
        for k, v in (env or {}).items():
            os.environ[k] = v
        try:
            forc = [to_layout(f, layout, torch.float32 if f32in else None) for f in plain64] if (layout == "tiled" or f32in) else plain64
            b = LandBatch(store, cal, list(synth.PERLND_FORCINGS), save, device=0, layout=layout, jit=jit)
            b.jit_word = word
            if f32in:
                b.set_forcing_dtype(np.float32)
            if f32out:
                b.set_output_dtype(np.float32)
            if sub is not None:
                b.set_schedule(sub)
            outs = B.out_buffers(torch, npad, Tc, b.ns, dev, layout, torch.float32 if f32out else None)
            for c in range(W):
                b.run_device(c * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
            evs[0].record(stream)
            for c in range(W, W + K):
                b.run_device(c * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
                evs[c - W + 1].record(stream)
            torch.cuda.synchronize()
            per = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
            ms = sum(per)
            got = outs[(W + K - 1) % 2].cpu().numpy()
            col = np.stack([B.out_column(got, i, layout) for i in (0, n // 2, n - 1)])
            key = "f32" if f32out else "f64"
            same = None
            if key in ref_out:
                same = bool(np.array_equal(col, ref_out[key], equal_nan=True))
            elif not f32in:
                ref_out[key] = col
            byt = (4 if f32in else 8) * 6 + (4 if f32out else 8) * 24
            rec = {"exp": "device", "label": label, "kernel": b.kernel_name(), "jit_error": b.jit_error,
                   "regs": (b.jit_meta or {}).get("registers"), "ms_per_step": ms / K, "value": n * Tc * K / (ms * 1e-3),
                   "bytes_per_segstep": byt, "frac_of_hbm": byt * n * Tc * K / (ms * 1e-3) / 1e9 / peak,
                   "per_step_ms": [round(x, 3) for x in per], "same_bits_as_first": same}
            print(json.dumps(rec), flush=True)
            b.close()
            del outs
            if forc is not plain64:
                del forc
            torch.cuda.empty_cache()
        except Exception as ex:
            print(json.dumps({"exp": "device", "label": label, "error": repr(ex)}), flush=True)
        finally:
            for k in (env or {}):
                os.environ.pop(k, None)

    run("tiled, library kernel (StaticIO, run-time option words)", "tiled", False)
    run("tiled, built for the batch", "tiled", True)
    run("plain, built for the batch, tensor-map TMA", "plain", True)
    run("plain, built for the batch, row-wise bulk copies", "plain", True, env={"HSP2G_NO_TMAP": "1"})
    run("plain, built for the batch, per-segment option words", "plain", True, word=False)
    if not os.environ.get("EXP_SHORT"):
        run("plain, library kernel (DynIO)", "plain", False)
    run("plain, float32 forcings", "plain", True, f32in=True)
    run("plain, float32 forcings and outputs", "plain", True, f32in=True, f32out=True)
    run("tiled, float32 outputs", "tiled", True, f32out=True)
    run("plain, float32 outputs", "plain", True, f32out=True)
    for sub in (() if os.environ.get("EXP_SHORT") else (32, 128)):
        run("plain, sub-chunk %d" % sub, "plain", True, sub=sub)
    del plain64
    torch.cuda.empty_cache()
    if os.environ.get("EXP_SKIP_E2E"):
        return

    # ---- end to end from plain host buffers through LandBatch.run_chunk (hsp2g_run_host_ld)
    def e2e(label, f32in, f32out, wc, steps=12, pageable=False, registered=False):
        try:
            b = LandBatch(store, cal, list(synth.PERLND_FORCINGS), save, device=0, layout="plain", jit=True)
            fd = np.float32 if f32in else np.float64
            od = np.float32 if f32out else np.float64
            b.set_forcing_dtype(fd)
            b.set_output_dtype(od)
            R = 3
            if pageable or registered:
                hf = [np.empty((Tc, b.nf, n), dtype=fd) for _ in range(R)]
                ho = [np.empty((Tc, b.ns, n), dtype=od) for _ in range(2)]
            else:
                hf = [pinned_empty((Tc, b.nf, n), fd, write_combined=wc) for _ in range(R)]
                ho = [pinned_empty((Tc, b.ns, n), od) for _ in range(2)]
            for j in range(R):
                hf[j][...] = synth.forcing_chunk(ens, cal, j * Tc, Tc, names=b.forcing_rows)
            t_reg = 0.0
            if registered:
                t0 = time.perf_counter()
                for a in hf + ho:
                    _lib.check(b.lib.hsp2g_host_register(a.ctypes.data, a.nbytes))
                t_reg = time.perf_counter() - t0
            for c in range(2):
                b.run_chunk(c * Tc, hf[c % R], ho[c % 2])
            b.sync()
            t0 = time.perf_counter()
            for c in range(2, 2 + steps):
                b.run_chunk(c * Tc, hf[c % R], ho[c % 2])
            b.sync()
            dt = time.perf_counter() - t0
            chk = float(np.nansum(ho[(1 + steps) % 2][-1, 13:], dtype=np.float64))
            if registered:
                for a in hf + ho:
                    b.lib.hsp2g_host_unregister(a.ctypes.data)
            hb, ob = hf[0].nbytes, ho[0].nbytes
            print(json.dumps({"exp": "e2e", "label": label, "kernel": b.kernel_name(), "value": n * Tc * steps / dt,
                              "ms_per_step": 1e3 * dt / steps, "h2d_bytes": hb, "d2h_bytes": ob,
                              "pcie_gbs": (hb + ob) * steps / dt / 1e9, "register_s": round(t_reg, 3), "checksum": chk}), flush=True)
            b.close()
            del hf, ho
        except Exception as ex:
            print(json.dumps({"exp": "e2e", "label": label, "error": repr(ex)}), flush=True)

    e2e("pinned, fp64 forcings, float32 outputs", False, True, False)
    e2e("pinned write-combined forcings, fp64 in, float32 out", False, True, True)
    e2e("pinned, float32 forcings, float32 outputs", True, True, False)
    e2e("pinned write-combined, float32 in / out", True, True, True)
    e2e("pinned, fp64 in / out", False, False, False, steps=6)
    e2e("registered numpy arrays (cudaHostRegister), fp64 in, float32 out", False, True, False, registered=True)
    e2e("pageable numpy arrays, fp64 in, float32 out", False, True, False, steps=4, pageable=True)


if __name__ == "__main__":
    main()
