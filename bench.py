#!/usr/bin/env python
"""bench.py -- segment-timesteps/s of the fused PERLND SNOW+PWATER kernel (fp64) on B200, with roofline and CPU baseline.

Workload (BASELINE.json configs[1]): test10 PERLND 1, SNOW+PWATER, replicated to 100,000 segments per GPU with jittered
parameters, 10-year hourly synthetic met, default SAVE set (13 SNOW + 11 PWATER series).
One "step" = one time chunk of `--chunk` intervals (default 438: the default 20 timed steps are exactly one simulated year,
so the number averages the cheap summer intervals and the expensive snow / melt ones) over all segments of the rank; steps
walk forward through simulated time, state carried.  Forcings of every timed step are resident in HBM before the timed region (each chunk is ~2.5 GB,
far larger than the 126 MB L2, so nothing is served from cache between steps).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, segments sharded, no collective)

`--impl reference` times the CPU oracle port (oracle/hsp2_oracle.c, OpenMP over segments) on the host cores: the
reference itself is pure Python + Numba and cannot travel to the GPU box.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# before the first CUDA call of the process (hspsquared_b200/__init__.py does the same): one hardware queue per stream for the
# ensembles that run one kernel per option word (engine.GroupedLandBatch)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

BYTES_PER_SEGSTEP = 8 * (6 + 24)  # 6 forcings read + 24 default-SAVE series written, fp64
N_SEG_PER_GPU = 100_000
RUN_START, RUN_STOP = "1976-01-01", "1986-01-01"  # 87,672 hourly intervals


def default_save():
    from hspsquared_b200 import test10

    return ["SNOW_" + n for n in test10.SAVE_DEFAULT["SNOW"]] + ["PWATER_" + n for n in test10.SAVE_DEFAULT["PWATER"]]


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe), polled through NVML in a
    thread (the timed region is tens of milliseconds: an nvidia-smi subprocess would not produce a sample in time);
    falls back to `nvidia-smi -lms` when the NVML binding is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.samples = []  # (sm_mhz, reasons bitmask)

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, r))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
            n = self.nvml
            names = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                     ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                     ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                     ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
            reasons = sorted({nm for _mhz, r in self.samples for nm, bit in names if r & bit})
            sm = [m for m, _r in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "NVML"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def measured_traffic(kernel, n, Tc):
    """DRAM bytes per launch of this kernel at this size from the committed ncu --set full capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            for e in json.load(f):
                if e["kernel"] == kernel and e["segments"] == n and e["intervals"] == Tc:
                    return e["dram_bytes_per_launch"], e["source"]
    except (OSError, ValueError, KeyError):
        pass
    return None, None


# ----------------------------------------------------------------------------------------------------------------
def cpu_sample(nseg, nsteps, seed=1234):
    """Inputs of the CPU baseline in the oracle's segment-major layout, from the same synthetic generator."""
    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import SEASONS, Calendar
    from oracle import oracle

    cal = Calendar(np.datetime64(RUN_START), np.datetime64(RUN_START) + np.timedelta64(nsteps, "h"), 60, "ns")
    ens = synth.Ensemble("PERLND", nseg, seed=seed, sections=("SNOW", "PWATER"))
    f = synth.forcing_chunk(ens, cal, 0, nsteps)  # [t][6][n]
    forc = np.ascontiguousarray(f.transpose(2, 1, 0))  # [n][6][t]
    names = oracle.fields("batch_par")
    par = np.zeros((nseg, len(names)))
    m0, m1, xm, dxm, _ = cal.day_interp()
    mon = np.zeros((nseg, 4, nsteps + 1))
    for i in range(nseg):
        t = ens.tables_of(i)
        flat = {}
        for sec in ("SNOW", "PWATER"):
            for grp in ("FLAGS", "PARAMETERS", "STATES"):
                flat.update(t[sec].get(grp, {}))
        for j, nm in enumerate(names):
            par[i, j] = float(flat[nm])
        for r, nm in enumerate(("CEPSC", "UZSN", "NSUR", "LZETP")):
            tab = np.array(list(t["PWATER"]["MONTHLY_" + nm].values()))
            v = ((tab[m1] - tab[m0]) / dxm) * xm + tab[m0]
            mon[i, r] = np.where(xm == 0, tab[m0], v)
    calarr = np.stack([cal.hoursval(np.ones(24), dofirst=True), cal.hourflag(0, dofirst=True),
                       cal.hour6flag(dofirst=True), cal.monthval(SEASONS)])
    return forc, par, mon, calarr


def host_threads():
    """Cores this process may run on (torchrun exports OMP_NUM_THREADS=1, so the count is passed explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu(nthreads, nseg, nsteps, min_seconds, max_reps=5000):
    from oracle import oracle

    forc, par, mon, cal = cpu_sample(nseg, nsteps)
    oracle.perlnd_batch(nsteps, 60.0, forc[: max(1, nseg // 8)], par[: max(1, nseg // 8)], mon[: max(1, nseg // 8)], cal,
                        nthreads)  # warm-up
    reps, t_total, used = 0, 0.0, nthreads
    while reps < max_reps and (t_total < min_seconds or reps < 1):
        t0 = time.perf_counter()
        sums, errs, used = oracle.perlnd_batch(nsteps, 60.0, forc, par, mon, cal, nthreads)
        t_total += time.perf_counter() - t0
        reps += 1
    assert np.isfinite(sums[:, 13:]).all()
    return nseg * nsteps * reps / t_total, used, reps, t_total


WORKLOAD = ("test10 PERLND SNOW+PWATER replicated to %d segments/GPU, jittered params, 10-yr hourly synthetic met, "
            "default SAVE (13 SNOW + 11 PWATER)")


def numba_reference(procs, per_proc, serial_segments, repeats=1, warm=1, timeout=900):
    """The UNMODIFIED reference's Numba implementation of the path (installed under baseline/_ref), serial and in a
    multiprocessing pool, in a subprocess of its own (oracle/numba_baseline.py) -> its JSON record."""
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/hsp2_ref_numba_cache")
    cmd = [sys.executable, "-W", "ignore", "-m", "oracle.numba_baseline", "--procs", str(procs), "--segments-per-proc", str(per_proc),
           "--serial-segments", str(serial_segments), "--repeats", str(repeats), "--warm", str(warm)]
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
        for line in reversed(r.stdout.splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": "oracle.numba_baseline produced no record: " + (r.stderr or "")[-300:]}
    except Exception as ex:
        return {"unavailable": repr(ex)}


def port_baseline(cores, budget_s=8.0):
    """The C restatement of the same cores (oracle/hsp2_oracle.c), OpenMP over segments: a second, faster CPU figure"""
    from oracle import oracle

    oracle.build()
    nsteps = 8784
    nseg = min(4096, 64 * cores)
    v, used, reps, tt = time_cpu(cores, nseg, nsteps, budget_s)
    vs, _, _, _ = time_cpu(1, 48, nsteps, 2.0)
    return {"value": v, "unit": "segment-timesteps/s", "cores": used, "kind": "port",
            "sample": "%d segments x %d hourly intervals (1976), SNOW+PWATER, x%d repeats, %.1f s; OpenMP over segments"
                      % (nseg, nsteps, reps, tt), "serial_value": vs}


def cpu_baseline():
    """The CPU implementations of the path on this box's host cores, beside the GPU number: (1) the reference itself -- its
    Numba kernels through its own snow() / pwater(), serial and multiprocess over segments (kind "reference"; kernel-only =
    time inside the @njit cores; wrapper-inclusive = what a call of the reference costs with its pandas glue);
    (2) the C port with OpenMP.  A reported baseline, not the optimisation target."""
    cores = host_threads()
    ref = numba_reference(cores, 12, 16)
    try:
        port = port_baseline(cores)
    except Exception as ex:
        port = {"error": repr(ex)}
    if "multiprocess" in ref:
        m = ref["multiprocess"]
        return {"value": m["kernel_only"], "unit": "segment-timesteps/s", "cores": m["processes"], "kind": "reference",
                "sample": "%d segments x %d hourly intervals (1976) of the bench ensemble in a multiprocessing.Pool(%d): time inside "
                          "the reference's @njit cores _snow_ + _pwater_ (kernel-only, slowest process)" % (m["segments"], m["intervals"], m["processes"]),
                "wrapper_inclusive_value": m["wrapper_inclusive"], "serial_value": ref["serial"]["kernel_only"],
                "serial_wrapper_inclusive_value": ref["serial"]["wrapper_inclusive"], "serial_cores": 1, "host_cpus": cores,
                "implementation": ref["implementation"], "numba": ref.get("numba"), "loaded_from": ref.get("reference_loaded_from"),
                "port": port}
    port["reference_unavailable"] = ref.get("unavailable")
    port["host_cpus"] = cores
    return port


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores -- the unmodified respec/HSPsquared
    package from baseline/_ref, its snow() and pwater() for segments of the same synthetic ensemble in a multiprocessing pool over
    all cores.  One step = one pass of the pool over a bounded sample (a few dozen segment-years per process).  value = kernel-only
    rate (time inside the @njit cores); the wrapper-inclusive rate and the C port's are in the same line."""
    if rank != 0:
        return
    cores = host_threads()
    per_proc = 32
    ref = numba_reference(cores, per_proc, 8, repeats=args.steps, warm=max(1, args.warmup))
    n = N_SEG_PER_GPU
    config = {"workload": WORKLOAD % n, "segments_per_gpu": n, "intervals_per_step": 438, "forcings": 6, "saved_series": 24}
    if "multiprocess" in ref:
        m = ref["multiprocess"]
        v = m["kernel_only"]
        sample = ("bounded sample of the configuration: %d segments x %d hourly intervals (one year) per step in a "
                  "multiprocessing.Pool(%d); time inside the reference's @njit cores, slowest process" % (m["segments"], m["intervals"], m["processes"]))
        cb = {"value": v, "unit": "segment-timesteps/s", "cores": m["processes"], "kind": "reference", "sample": sample,
              "wrapper_inclusive_value": m["wrapper_inclusive"], "serial_value": ref["serial"]["kernel_only"],
              "implementation": ref["implementation"], "numba": ref.get("numba"), "loaded_from": ref.get("reference_loaded_from")}
        ms = 1e3 * m["wall_s_per_pass"]
    else:  # no installed reference on this box: the C port stands in (kind "port")
        cb = port_baseline(cores, budget_s=max(4.0, 0.5 * args.steps))
        cb["reference_unavailable"] = ref.get("unavailable")
        v, ms = cb["value"], None
    line = {
        "impl": "reference", "metric": "segment-timesteps/sec PERLND SNOW+PWATER fp64", "value": v,
        "unit": "segment-timesteps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "segment-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        line["c_port_openmp"] = port_baseline(cores, budget_s=4.0) if "multiprocess" in ref else None
    except Exception as ex:
        line["c_port_openmp"] = {"error": repr(ex)}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
def device_forcing(torch, ens, cal, t0, nsteps, npad, dev, rows, layout="plain", dtype=None):
    """fp64 forcings of intervals [t0, t0+nsteps) generated ON the device with the same elementwise expressions as
    synth.forcing_chunk (separate mul / add kernels: no FMA), so the oracle spot-check sees identical bits.
    layout "plain": [nsteps][rows][npad] (include/hsp2g.h HSP2G_PLAIN, ld = npad); "tiled": [ntiles][nsteps][rows][32].
    Columns past n copy the last segment."""
    from hspsquared_b200 import synth

    s = synth.shared_series(ens.operation, cal, t0, nsteps)
    n = ens.n
    f = torch.zeros((nsteps, len(rows), npad), dtype=torch.float64, device=dev)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    airt = T(s["AIRTMP"])[:, None] + T(3.0 * ens.u_airt)[None, :]
    for r, nm in enumerate(rows):
        if nm == "PREC":
            f[:, r, :n] = T(s["PREC"])[:, None] * T(1.0 + 0.2 * ens.u_prec)[None, :]
        elif nm == "PETINP":
            f[:, r, :n] = T(s["PETINP"])[:, None] * T(1.0 + 0.1 * ens.u_pet)[None, :]
        elif nm in ("AIRTMP", "GATMP"):
            f[:, r, :n] = airt
        elif nm == "DTMPG":
            f[:, r, :n] = airt - 5.0
        else:
            f[:, r, :n] = T(s[nm])[:, None]
    if npad > n:
        f[:, :, n:] = f[:, :, n - 1:n]
    if dtype is not None:
        f = f.to(dtype)
    if layout == "plain":
        return f
    return f.view(nsteps, len(rows), npad // 32, 32).permute(2, 0, 1, 3).contiguous()


def out_buffers(torch, npad, Tc, ns, dev, layout="plain", dtype=None, count=2):
    shape = (Tc, ns, npad) if layout == "plain" else (npad // 32, Tc, ns, 32)
    return [torch.empty(shape, dtype=dtype or torch.float64, device=dev) for _ in range(count)]


def out_column(got, i, layout):
    """series [interval][row] of segment i from a device output buffer brought to the host"""
    return got[:, :, i] if layout == "plain" else got[i // 32, :, :, i % 32]


OVERLAP = os.environ.get("HSP2G_OVERLAP", "1") != "0"  # measurement switch: 0 = every launch waits for the previous one to end


def time_device(torch, batch, forc, outs, Tc, K, W, stream, dev, dist, first=0):
    """W untimed + K timed launches of `batch` over consecutive chunks (state carried), CUDA events on the launching stream,
    barrier + synchronize on both sides, max over ranks.  -> ms for the K launches"""

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # consecutive launches overlap at their tails (hsp2g_set_launch_overlap): every chunk's forcings are resident before the
    # loop starts and the two output buffers alternate, which is the caller's side of that contract
    overlap = OVERLAP and hasattr(batch, "set_launch_overlap")
    if overlap:
        batch.set_launch_overlap(True)
    for c in range(first, first + W):
        batch.run_device(c * Tc, Tc, forc[c % len(forc)].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for c in range(first + W, first + W + K):
        batch.run_device(c * Tc, Tc, forc[c % len(forc)].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if overlap:
        batch.set_launch_overlap(False)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--segments", type=int, default=N_SEG_PER_GPU, help="segments per GPU")
    ap.add_argument("--chunk", type=int, default=438,
                    help="intervals per step; the default makes the default 20 timed steps exactly one simulated year "
                         "(8,760 hourly intervals), so the figure is an annual average over snow, melt and summer regimes")
    ap.add_argument("--layout", default="plain", choices=("plain", "tiled"),
                    help="series layout in HBM: plain = [interval][row][segment] (the reference's orientation, tensor-map TMA), "
                         "tiled = [tile][interval][row][32]")
    ap.add_argument("--sub", type=int, default=None, help="intervals per (tile, sub-chunk) task; default = library's")
    ap.add_argument("--no-jit", action="store_true", help="run the library's general kernels instead of one built for the batch")
    ap.add_argument("--caller-order", action="store_true", help="keep the ensemble generator's segment order on the device "
                    "instead of the engine's automatic divergence-aware order")
    ap.add_argument("--no-ordered", action="store_true", help="skip the extra measurement in the other segment order")
    ap.add_argument("--no-overlap", action="store_true", help="every launch waits for the previous one to end (default: consecutive "
                    "launches of a batch overlap at their tails, hsp2g_set_launch_overlap)")
    ap.add_argument("--no-alt-layout", action="store_true", help="skip the extra measurement in the other series layout")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations timed beside the headline")
    ap.add_argument("--no-ensemble", action="store_true", help="N > 1: skip the sharded 1M-segment PERLND+IMPLND ensemble (configs[2])")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.no_overlap:
        global OVERLAP
        OVERLAP = False

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the land-segment path has no CPU fallback (use --impl reference for the CPU port)")
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"  # keep stdout to the one JSON line (at VERSION and above NCCL prints a banner there)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    pin_to_gpu_numa_node(local_rank)

    from hspsquared_b200 import _lib, synth
    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import LandBatch

    _lib.load()
    n, Tc, K, W = args.segments, args.chunk, args.steps, args.warmup
    cal = Calendar(RUN_START, RUN_STOP, 60, "ns")
    total_chunks = K + W
    if total_chunks * Tc > cal.steps:
        sys.exit("steps*chunk exceeds the 10-year run (%d intervals)" % cal.steps)
    ens0 = synth.Ensemble("PERLND", n, seed=1234 + rank, sections=("SNOW", "PWATER"))
    # Segment order on the device: the engine's divergence-aware ordering, automatic -- option word, then a climate key read off
    # the first week of the batch's own forcings (LandBatch.climate_order).  Same segments, same arithmetic, same bits per
    # segment; `--caller-order` keeps the generator's order (also measured beside the headline unless --no-ordered).
    if args.caller_order:
        ens = ens0
    else:
        sample = synth.forcing_chunk(ens0, cal, 0, 24 * 7)
        ens = ens0.take(LandBatch.climate_order(ens0.store(cal), list(synth.PERLND_FORCINGS), sample))
        del sample
    store = ens.store(cal)
    save = default_save()
    jit = not args.no_jit
    batch = LandBatch(store, cal, list(synth.PERLND_FORCINGS), save, device=local_rank, layout=args.layout, jit=jit, order=None)
    if args.sub is not None:
        batch.set_schedule(args.sub)
    npad, nf, ns = batch.npad, batch.nf, batch.ns
    assert nf == 6 and ns == 24

    # ---- device-resident inputs for every step of the run, two output slots
    free, _tot = torch.cuda.mem_get_info(dev)
    per_f = Tc * nf * npad * 8
    ring = max(2, min(total_chunks, int((free * 0.6 - 2 * Tc * ns * npad * 8) // per_f)))
    forc = [device_forcing(torch, ens, cal, c * Tc, Tc, npad, dev, batch.forcing_rows, args.layout) for c in range(min(ring, total_chunks))]
    outs = out_buffers(torch, npad, Tc, ns, dev, args.layout)
    torch.cuda.synchronize(dev)
    stream = torch.cuda.Stream(dev)  # a real (non-default) stream: the kernels, and the events that time them, live here
    assert stream.cuda_stream != 0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    l0, _ = batch.kernel_stats()
    sampler.start()  # the warm-up launches are sampled too; the timed region is tens of milliseconds
    ms = time_device(torch, batch, forc, outs, Tc, K, W, stream, dev, dist)
    clocks = sampler.stop()
    l1, _ = batch.kernel_stats()
    kname = batch.kernel_name()
    n_forcing_chunks = len(forc)
    seg_steps = n * Tc * K * world
    value = seg_steps / (ms * 1e-3)

    # ---- parity spot-check of the timed run against the CPU oracle (rank 0): last chunk, three segments
    parity = None
    if rank == 0:
        try:
            from oracle import cases as CASES
            from oracle import wrappers as Wr

            last = W + K - 1
            got = outs[last % 2].cpu().numpy()
            nst = (last + 1) * Tc
            acts = {"SNOW": Wr.snow, "PWATER": Wr.pwater}
            worst = 0.0
            for i in (0, n // 2, n - 1):
                si = {"start": cal.start, "stop": cal.start + np.timedelta64(nst, "h"), "delt": 60, "units": 1,
                      "steps": nst, "time_unit": "ns"}
                ts = synth.forcing_of(ens, cal, i, 0, nst)
                CASES.run_segment("PERLND", ens.tables_of(i), ts, si, acts)
                col = out_column(got, i, args.layout)
                for r, full in enumerate(batch.save_rows):
                    ref = ts[full.split("_", 1)[1]][last * Tc:]
                    g = col[:, r]
                    ok = ~np.isnan(ref)
                    assert np.array_equal(np.isnan(g), ~ok)
                    den = np.maximum(np.abs(ref[ok]), max(1e-4 * np.abs(ref[ok]).max(), 1e-300)) if ok.any() else 1.0
                    if ok.any():
                        worst = max(worst, float((np.abs(g[ok] - ref[ok]) / den).max()))
            parity = {"max_rel_err_vs_oracle": worst, "bit_identical": worst == 0.0, "segments_checked": 3, "intervals": nst,
                      "tolerance": 1e-9}
        except Exception as ex:  # the number stands; say that the check could not run
            parity = {"error": repr(ex)}

    peak, peak_src = measured_peak()
    frac_of = lambda ms_: BYTES_PER_SEGSTEP * n * Tc / (ms_ * 1e-3 / K) / 1e9 / peak

    # ---- the same year in the other series layout (same kernel template, same bits): reported beside the headline
    alt = None
    if not args.no_alt_layout:
        other = "tiled" if args.layout == "plain" else "plain"
        try:
            b2 = LandBatch(store, cal, list(synth.PERLND_FORCINGS), save, device=local_rank, layout=other, jit=jit)
            if args.sub is not None:
                b2.set_schedule(args.sub)
            for c in range(len(forc)):  # re-layout chunk by chunk, in place of the first copy
                f = forc[c]
                forc[c] = None
                forc[c] = (f.view(Tc, nf, npad // 32, 32).permute(2, 0, 1, 3).contiguous() if other == "tiled"
                           else f.permute(1, 2, 0, 3).reshape(Tc, nf, npad).contiguous())
                del f
            outs2 = out_buffers(torch, npad, Tc, ns, dev, other)
            ms2 = time_device(torch, b2, forc, outs2, Tc, K, W, stream, dev, dist)
            same = bool(torch.equal(torch.nan_to_num(outs2[(W + K - 1) % 2].cpu() if other == "plain" else
                                                     outs2[(W + K - 1) % 2].permute(1, 2, 0, 3).reshape(Tc, ns, npad).cpu()),
                                    torch.nan_to_num(outs[(W + K - 1) % 2].cpu() if args.layout == "plain" else
                                                     outs[(W + K - 1) % 2].permute(1, 2, 0, 3).reshape(Tc, ns, npad).cpu())))
            alt = {"layout": other, "value": seg_steps / (ms2 * 1e-3), "ms_per_step": ms2 / K, "frac": frac_of(ms2),
                   "kernel": b2.kernel_name(), "same_bits_as_headline": same}
            b2.close()
            del outs2
        except Exception as ex:
            alt = {"error": repr(ex)}

    # ---- the same year with the segments in the other order (the generator's own / the automatic one): beside the headline
    ordered = None
    if not args.no_ordered:
        try:
            del forc
            torch.cuda.empty_cache()
            if args.caller_order:
                sample = synth.forcing_chunk(ens0, cal, 0, 24 * 7)
                ens2 = ens0.take(LandBatch.climate_order(ens0.store(cal), list(synth.PERLND_FORCINGS), sample))
                del sample
                what = "automatic divergence-aware order (LandBatch.climate_order)"
            else:
                ens2, what = ens0, "segments in the ensemble generator's own order (no ordering)"
            b2 = LandBatch(ens2.store(cal), cal, list(synth.PERLND_FORCINGS), save, device=local_rank, layout=args.layout, jit=jit,
                           order=None)
            if args.sub is not None:
                b2.set_schedule(args.sub)
            forc2 = [device_forcing(torch, ens2, cal, c * Tc, Tc, npad, dev, b2.forcing_rows, args.layout) for c in range(total_chunks)]
            torch.cuda.synchronize(dev)
            ms2 = time_device(torch, b2, forc2, outs, Tc, K, W, stream, dev, dist)
            b2.close()
            del forc2
            torch.cuda.empty_cache()
            ordered = {"order": what, "value": seg_steps / (ms2 * 1e-3), "ms_per_step": ms2 / K, "frac": frac_of(ms2)}
        except Exception as ex:
            ordered = {"error": repr(ex)}
    forc = None
    del outs
    torch.cuda.empty_cache()

    traffic, traffic_src = measured_traffic(kname, n, Tc)
    avg_launch_s = ms * 1e-3 / K
    achieved = BYTES_PER_SEGSTEP * n * Tc / avg_launch_s / 1e9
    line = {
        "metric": "segment-timesteps/sec PERLND SNOW+PWATER fp64", "value": value, "unit": "segment-timesteps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % n,
                   "segments_per_gpu": n, "intervals_per_step": Tc, "forcings": 6, "saved_series": 24,
                   "series_layout": "%s (%s)" % (args.layout, "[interval][row][segment], as the reference's ts arrays"
                                                 if args.layout == "plain" else "[tile][interval][row][32]"),
                   "segment_order": "caller's" if args.caller_order else "automatic: option word, then climate key from the first "
                                    "week of the batch's forcings (LandBatch.climate_order)",
                   "cache": "inputs of every step (%.2f GB) and outputs (%.2f GB) are far larger than L2; %d distinct "
                            "forcing chunks resident" % (per_f / 1e9, Tc * ns * npad * 8 / 1e9, n_forcing_chunks),
                   "launches": ("consecutive launches overlap at their tails (programmatic dependent launch, per-tile hand-over; "
                                "hsp2g_set_launch_overlap)" if OVERLAP else "each launch waits for the previous one to end"),
                   "kernel": kname, "parallelism": "segments sharded over %d GPU(s), no collective" % world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_SEGSTEP * n * Tc,
                     "avg_launch_ms": avg_launch_s * 1e3},
        "gpu_launches": int(l1 - l0) - W,
        "other_layout": alt,
        "other_segment_order": ordered,
        "clocks": clocks,
        "parity": parity,
    }
    # the headline is measured: whatever happens in the legs below, rank 0 prints this line (a leg that does not come back within
    # HSP2G_BENCH_LEGS_S seconds is reported as such instead of taking the whole record with it)
    finished = threading.Event()

    def watchdog():
        limit = float(os.environ.get("HSP2G_BENCH_LEGS_S", "600"))
        if finished.wait(limit):
            return
        if rank == 0:
            part = dict(line)
            part["watchdog"] = "the legs after the headline (%s) did not finish within %.0f s; record printed without them" % (
                ", ".join(k for k in ("e2e", "ensemble_1m_sharded", "other_configs", "cpu_baseline") if k not in part) or "teardown", limit)
            sys.stdout.write(json.dumps(part) + "\n")
            sys.stdout.flush()
        os._exit(0)

    threading.Thread(target=watchdog, daemon=True).start()
    # ---- end to end through the public API with plain HOST buffers (H2D and D2H inside the timed region)
    if not args.no_e2e:
        line["e2e"] = run_e2e(torch, batch, ens, cal, store, n, Tc, K, W, dev, dist, world, barrier, ens0=ens0)
    batch.close()
    if world > 1 and not args.no_ensemble:
        try:
            r = run_sharded_ensemble(torch, dev, local_rank, rank, world, dist, cal, stream, peak, jit)
            if rank == 0:
                line["ensemble_1m_sharded"] = r
        except Exception as ex:
            line["ensemble_1m_sharded"] = {"error": repr(ex)}
    if not args.no_extras:
        try:
            line["other_configs"] = run_extras(torch, dev, local_rank, rank, world, dist, cal, stream, peak)
        except Exception as ex:
            line["other_configs"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline()
        except Exception as ex:
            line["cpu_baseline"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line))
        sys.stdout.flush()
    finished.set()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def gpu_local_cpus(index):
    """cores of the NUMA node / root complex GPU `index` hangs off, from sysfs; None where it does not say"""
    try:
        import torch

        p = torch.cuda.get_device_properties(index)
        addr = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % addr) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        return sorted(cpus) or None
    except Exception:
        return None


def pin_to_gpu_numa_node(local_rank):
    """Host threads of this rank run on the cores of its GPU's NUMA node, so that the pinned buffers it allocates (first touch)
    and the copies it drives stay local to the GPU's root complex.  A no-op where sysfs does not say (single-node VMs)."""
    try:
        cpus = gpu_local_cpus(local_rank)
        have = set(os.sched_getaffinity(0))
        allowed = sorted(set(cpus or ()) & have)
        if allowed and len(allowed) < len(have):
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def run_e2e(torch, batch, ens, cal, store, n, Tc, K, W, dev, dist, world, barrier, ens0=None):
    """Same metric through the public API a host program calls -- LandBatch.run_chunk -> hsp2g_run_host_ld -- with PLAIN host
    buffers, [interval][row][segment] exactly as the reference's ts arrays stack: nothing is re-tiled on the host, each chunk
    is one (strided) copy each way, H2D, kernel and D2H pipelined over 3 streams / 2 device slots, all inside the timed region.
    Outputs cross the bus as float32 -- the element type the reference's save_timeseries hands its IOManager
    (utilities.py:313).  Variants timed beside the headline: float32 forcings (exact for float32 sources such as WDM / HDF5),
    float64 outputs, daily statistics computed on the device, and pageable (ordinary numpy) buffers through LandBatch.run()."""
    from hspsquared_b200 import _lib, synth
    from hspsquared_b200.engine import LandBatch, pinned_empty

    lib = _lib.load()
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    nf, ns = batch.nf, batch.ns
    Te = Tc
    R = 3  # distinct forcing chunks resident on the host, sent round robin
    while Te > 32 and (R * nf * 8 + 2 * ns * 8) * Te * n > 0.35 * avail / world:  # every rank pins its own buffers
        Te //= 2
    if batch.layout != "plain":
        batch = LandBatch(store, cal, list(synth.PERLND_FORCINGS), list(batch.save), device=batch.device, layout="plain", jit=batch.jit)
        own = True
    else:
        own = False

    def timed(fdtype, odtype, steps, base, daily=False, wc=True):
        batch.set_forcing_dtype(fdtype)
        batch.set_output_dtype(odtype)
        hf = [pinned_empty((Te, nf, n), fdtype, write_combined=wc) for _ in range(R)]
        for j in range(R):
            hf[j][...] = synth.forcing_chunk(ens, cal, base + j * Te, Te, names=batch.forcing_rows)
        if daily:  # 24-interval statistics (last, sum, mean) computed on the device, a new period at every chunk start
            nb = (Te + 23) // 24
            t = np.arange(cal.steps, dtype=np.int64)
            p0 = (base - 1) // 24 if base else -1  # intervals before the timed region: plain 24-interval periods
            per = np.where(t < base, t // 24, p0 + 1 + ((t - base) // Te) * nb + ((t - base) % Te) // 24).astype(np.int32)
            per = np.ascontiguousarray(per)
            _lib.check(lib.hsp2g_set_output_periods(batch.h, len(per), per.ctypes.data_as(C.POINTER(C.c_int32))))
            ho = [pinned_empty((nb, ns, 3, n), np.float32) for _ in range(2)]
        else:
            ho = [pinned_empty((Te, ns, n), odtype) for _ in range(2)]
        nwarm = 2
        for c in range(nwarm):
            batch.run_host_ptr(base + c * Te, Te, hf[c % R].ctypes.data, ho[c % 2].ctypes.data)
        batch.sync()
        barrier()
        t0 = time.perf_counter()
        for c in range(nwarm, nwarm + steps):
            batch.run_host_ptr(base + c * Te, Te, hf[c % R].ctypes.data, ho[c % 2].ctypes.data)
        batch.sync()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        per_rank = None
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank = [round(1e3 * float(x.item()) / steps, 2) for x in allt]
            dt = max(float(x.item()) for x in allt)
        if daily:
            _lib.check(lib.hsp2g_set_output_periods(batch.h, 0, None))
        last = ho[(nwarm + steps - 1) % 2]
        chk = float(np.nansum(last[-1, 13:], dtype=np.float64))  # read the step's result
        fb, ob = hf[0].nbytes, ho[0].nbytes
        del hf, ho
        return {"value": n * Te * steps * world / dt, "ms_per_step": 1e3 * dt / steps, "h2d_bytes_per_step": fb,
                "d2h_bytes_per_step": ob, "pcie_gbs": (fb + ob) * steps / dt / 1e9, "result_checksum": chk, "steps": steps,
                "per_rank_ms_per_step": per_rank, "kernel": batch.kernel_name()}, base + (nwarm + steps) * Te

    def pageable(steps, base):
        """ordinary numpy arrays through LandBatch.run(): what a caller who knows nothing about pinned memory pays"""
        from hspsquared_b200.engine import LandBatch

        e0 = ens0 if ens0 is not None else ens  # the ensemble in the generator's own order: no ordering done for the caller
        f = synth.forcing_chunk(e0, cal, base, Te, names=list(synth.PERLND_FORCINGS))
        with LandBatch(e0.store(cal), cal, list(synth.PERLND_FORCINGS), default_save(), device=dev.index or 0) as plain_batch:
            return pageable_run(plain_batch, f, steps)

    def pageable_run(batch, f, steps):
        """a batch built with the defaults (the caller's segment order), ordinary arrays in, a float32 array out"""
        batch.set_output_dtype(np.float32)
        ck = (Te + 2) // 3  # three chunks per call: copy-in, kernel and copy-out of neighbouring chunks overlap
        batch.run(f, chunk=ck)  # warm-up
        barrier()
        t0 = time.perf_counter()
        chk = 0.0
        for _ in range(steps):
            o = batch.run(f, chunk=ck)
            chk = float(np.nansum(o[-1, 13:], dtype=np.float64))
        dt = time.perf_counter() - t0
        batch.run(f, chunk=ck, out=o)  # registers o
        t0 = time.perf_counter()
        for _ in range(steps):
            batch.run(f, chunk=ck, out=o)
        dt2 = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt, dt2], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt2 = float(t[0].item()), float(t[1].item())
        return {"value": n * Te * steps * world / dt, "ms_per_step": 1e3 * dt / steps, "steps": steps, "result_checksum": chk,
                "reused_output_array": {"value": n * Te * steps * world / dt2, "ms_per_step": 1e3 * dt2 / steps},
                "note": "ordinary numpy arrays in and out through LandBatch.run(): page-locked in place (cudaHostRegister) and kept "
                        "registered while the array lives; value: a NEW output array every call (first touch + registration of 4.2 GB "
                        "per step); reused_output_array: run(out=...) with the same arrays again -- nothing is registered twice"}

    try:
        # continue the simulation where the device-resident run stopped: keeps state and calendar consistent
        base = (W + K) * Tc
        K4 = max(2, K // 4)
        if base + (2 + K + 4 * (2 + K4)) * Te > cal.steps:
            base = 0
            batch.set_state(store.state)
        r32, base = timed(np.float64, np.float32, K, base)
        out = {"value": r32["value"], "unit": "segment-timesteps/s", "h2d_bytes_per_step": r32["h2d_bytes_per_step"],
               "d2h_bytes_per_step": r32["d2h_bytes_per_step"], "intervals_per_step": Te, "steps": K,
               "ms_per_step": r32["ms_per_step"], "pcie_gbs": r32["pcie_gbs"], "result_checksum": r32["result_checksum"],
               "per_rank_ms_per_step": r32["per_rank_ms_per_step"], "kernel": r32["kernel"],
               "host_buffers": "plain [interval][row][segment] pinned arrays (forcings write-combined), %d distinct forcing chunks "
                               "sent round robin; no host-side re-tiling" % R,
               "api": "LandBatch.run_host_ptr -> hsp2g_run_host_ld (one strided H2D, kernel, one strided D2H per step, pipelined "
                      "over 3 streams / 2 device slots)",
               "out_dtype": "float32 (the cast save_timeseries applies before IOManager.write_ts, utilities.py:313)",
               "forcing_dtype": "float64"}
        for key, fd, od, kw in (("float32_forcings", np.float32, np.float32, {}),
                                ("fp64_outputs", np.float64, np.float64, {}),
                                ("daily_statistics", np.float64, np.float32, {"daily": True}),
                                ("daily_statistics_float32_forcings", np.float32, np.float32, {"daily": True})):
            try:
                r, base = timed(fd, od, K4, base, **kw)
                out[key] = {k: r[k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "pcie_gbs", "steps",
                                             "kernel", "per_rank_ms_per_step")}
            except Exception as ex:
                out[key] = {"error": repr(ex)}
        out["float32_forcings"]["note"] = ("forcings cross PCIe and HBM as float32 and are widened in the kernel: exact when the "
                                           "source is float32 (WDM, HDF5 /TIMESERIES; utilities.py:283-287)")
        out["daily_statistics"]["note"] = ("hsp2g_set_output_periods: last / sum / mean per 24 intervals computed on the device "
                                           "(IOManager.write_ts outstep 3, hsp2io/io.py:74-81), only the statistics cross PCIe")
        if world > 2:  # 10 GB more of host arrays per rank: measured at N = 1 and 2 only
            out["pageable_numpy"] = {"skipped": "measured at N <= 2 (profiles/r3b_run_reuse_ordinary_arrays.json, r3c_bench_n2_line.json)"}
        else:
            try:
                out["pageable_numpy"] = pageable(max(2, K // 10), base)
            except Exception as ex:
                out["pageable_numpy"] = {"error": repr(ex)}
        return out
    except Exception as ex:
        return {"error": repr(ex)}
    finally:
        if own:
            batch.close()


CBP10 = ("PWATER_SURO", "PWATER_IFWO", "PWATER_AGWO", "SEDMNT_SOSED", "PWTGAS_SOHT", "PWTGAS_IOHT", "PWTGAS_AOHT",
         "PWTGAS_SODOXM", "PWTGAS_IODOXM", "PWTGAS_AODOXM")  # tests/land_spec/hwmA51800.uci:493-502
FULLCHAIN_OPTIONS = ("RTOPFG", "UZFG", "SNOPFG", "ICEFG", "TSOPFG")


def config_workload(name, n=None, seed=1234):
    """(ensemble, forcing names, SAVE set, segments, intervals per step, note) of a named BASELINE.json configuration"""
    from hspsquared_b200 import synth

    if name == "headline":  # configs[1]
        n = n or 100_000
        return synth.Ensemble("PERLND", n, seed=seed, sections=("SNOW", "PWATER")), list(synth.PERLND_FORCINGS), default_save(), n, 438, \
            "test10 PERLND SNOW+PWATER, default SAVE"
    if name == "implnd":  # the IMPLND half of configs[2]
        from hspsquared_b200 import test10

        n = n or 100_000
        save = ["SNOW_" + k for k in test10.SAVE_DEFAULT["SNOW"]] + ["IWATER_" + k for k in test10.SAVE_DEFAULT["IWATER"]]
        return synth.Ensemble("IMPLND", n, seed=seed, sections=("SNOW", "IWATER")), list(synth.IMPLND_FORCINGS), save, n, 438, \
            "test10 IMPLND SNOW+IWATER, default SAVE"
    if name in ("fullchain_all", "fullchain_cbp10", "fullchain_uniform_all", "fullchain_uniform_cbp10"):  # configs[3]
        n = n or 250_000
        uniform = "uniform" in name
        ens = synth.Ensemble("PERLND", n, seed=seed, base="cbp", options=() if uniform else FULLCHAIN_OPTIONS)
        save = "all" if name.endswith("_all") else list(CBP10)
        return ens, list(synth.FULLCHAIN_FORCINGS), save, n, 128 if save == "all" else 438, \
            "full PERLND chain ATEMP/SNOW/PWATER/SEDMNT/PSTEMP/PWTGAS on the CBP land segment (tests/land_spec/hwmA51800.uci), " + \
            ("one option word" if uniform else "option flags %s drawn per segment" % "/".join(FULLCHAIN_OPTIONS)) + \
            (", all 84 outputs" if save == "all" else ", the 10 series a CBP land segment hands its river model")
    raise KeyError(name)


def run_config(torch, dev, name, cal, stream, peak, n=None, layout="plain", jit=True, order="flags", K=20, W=3, Tc=None, sub=None,
               dist=None, out_dtype=None, first_chunk=0, seed=1234, forcing_dtype=None):
    """One device-resident measurement of a named configuration (same rules as the headline: W >= 3 warm-up launches, CUDA events
    on the launching stream, inputs of every launch far larger than L2 and distinct from the previous launch's).
    order: None | "flags" (option word) | "climate" (option word, then the climate key LandBatch.climate_order derives from the
    first chunk's forcings) | "groups" (one batch and kernel per option word, launched together: engine.GroupedLandBatch)."""
    from hspsquared_b200 import synth
    from hspsquared_b200.engine import LandBatch

    ens, forcings, save, n, Tc0, note = config_workload(name, n, seed)
    Tc = Tc or Tc0
    store = ens.store(cal)
    perm = None
    grouped = order == "groups"
    if order in ("climate", "groups"):
        f0 = synth.forcing_chunk(ens, cal, first_chunk * Tc, min(Tc, 24 * 7), names=forcings)
        if grouped:
            keys = LandBatch.climate_keys(forcings, f0)
        else:
            perm = LandBatch.climate_order(store, forcings, f0)
        del f0
    elif order == "flags":
        perm = store.divergence_order()
    if perm is not None and not np.array_equal(perm, np.arange(n)):
        ens = ens.take(perm)
        store = ens.store(cal)
    if grouped:
        # one batch and one kernel per option word, launched together (engine.GroupedLandBatch); climate keys inside a group
        from hspsquared_b200.engine import GroupedLandBatch

        batch = GroupedLandBatch(store, cal, forcings, save, device=dev.index or 0, layout=layout, jit=jit, keys=keys)
    else:
        batch = LandBatch(store, cal, forcings, save, device=dev.index or 0, layout=layout, jit=jit, order=None)
    try:
        if sub is not None:
            batch.set_schedule(sub)
        if out_dtype is not None:
            batch.set_output_dtype(out_dtype)
        f32in = forcing_dtype is not None and np.dtype(forcing_dtype).itemsize == 4 and not grouped
        if f32in:
            batch.set_forcing_dtype(np.float32)
        nf, ns = batch.nf, batch.ns
        osz = 4 if out_dtype is not None and np.dtype(out_dtype).itemsize == 4 else 8
        if grouped:
            if layout != "plain":
                raise ValueError("grouped runs of run_config use the plain layout")
            npad = batch.width
            parts = [(c0, b.npad, ens.take(idx)) for c0, idx, b in zip(batch.first_column, batch.members, batch.groups)]
            forc = []
            for c in range(K + W):
                big = torch.empty((Tc, nf, npad), dtype=torch.float64, device=dev)
                for c0, w_, e_ in parts:
                    big[:, :, c0:c0 + w_] = device_forcing(torch, e_, cal, (first_chunk + c) * Tc, Tc, w_, dev, batch.forcing_rows, "plain")
                forc.append(big)
        else:
            npad = batch.npad
            forc = [device_forcing(torch, ens, cal, (first_chunk + c) * Tc, Tc, npad, dev, batch.forcing_rows, layout,
                                   torch.float32 if f32in else None) for c in range(K + W)]
        outs = out_buffers(torch, npad, Tc, ns, dev, layout, torch.float32 if osz == 4 else None)
        torch.cuda.synchronize(dev)
        ms = time_device(torch, batch, forc, outs, Tc, K, W, stream, dev, dist, first=0) if first_chunk == 0 else None
        if ms is None:  # a window that does not start at interval 0: chunk indices are offset
            class _Shift(list):
                pass
            for c in range(W):
                batch.run_device((first_chunk + c) * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for c in range(W, W + K):
                batch.run_device((first_chunk + c) * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
        B = (4 if f32in else 8) * nf + osz * ns
        value = n * Tc * K / (ms * 1e-3)
        errs = {k: int((v != 0).sum()) for k, v in batch.errors().items() if (v != 0).any()}
        return {"config": name, "workload": note, "segments": n, "intervals_per_step": Tc, "steps": K, "warmup": W,
                "window": "intervals %d..%d of the run (from 1976-01-01: the snow and melt season, the expensive regime)"
                          % ((first_chunk + W) * Tc, (first_chunk + W + K) * Tc),
                "series_layout": layout, "segment_order": order,
                "kernel": batch.kernel_name() if not grouped else "%d kernels, one per option word, launched together; e.g. %s"
                % (len(batch.groups), batch.kernel_names()[0]),
                "registers": (batch.jit_meta or {}).get("registers") if not grouped else sorted({(b.jit_meta or {}).get("registers") for b in batch.groups}, key=str),
                "forcing_rows": nf, "saved_rows": ns,
                "bytes_per_segment_interval": B, "ms_per_step": ms / K, "value": value, "unit": "segment-timesteps/s",
                "achieved_GBps": value * B / 1e9, "frac_of_measured_hbm": value * B / 1e9 / peak, "segments_with_error_counts": errs}
    finally:
        batch.close()
        torch.cuda.empty_cache()


def run_sharded_ensemble(torch, dev, local_rank, rank, world, dist, cal, stream, peak, jit, n_total=1_000_000, Tc=128, K=12, W=3):
    """BASELINE.json configs[2]: ONE 1M-segment PERLND + IMPLND water-balance ensemble (500k + 500k) split over the ranks with
    dist.shard_bounds (hspsquared_b200.dist.ShardedBatch); every rank advances its shard of both operations, device-resident;
    no collective on the data path; at the end the per-segment sums of the runoff series are gathered to every rank over NCCL
    (the numbers a calibration run keeps per ensemble member).  Strong scaling: the total is fixed, per-GPU work shrinks with N."""
    import torch.distributed as tdist

    from hspsquared_b200 import synth, test10
    from hspsquared_b200.dist import ShardedBatch

    half = n_total // 2
    parts = []
    for op, secs, save in (("PERLND", ("SNOW", "PWATER"), default_save()),
                           ("IMPLND", ("SNOW", "IWATER"), ["SNOW_" + k for k in test10.SAVE_DEFAULT["SNOW"]] +
                            ["IWATER_" + k for k in test10.SAVE_DEFAULT["IWATER"]])):
        ens = synth.Ensemble(op, half, seed=777, sections=secs)
        sb = ShardedBatch(lambda lo, hi, e=ens: e.shard(lo, hi).store(cal), half, cal, list(synth.PERLND_FORCINGS), save,
                          device=local_rank, layout="plain", jit=jit, order=None)
        shard = ens.shard(sb.lo, sb.hi)
        b = sb.batch
        forc = [device_forcing(torch, shard, cal, c * Tc, Tc, b.npad, dev, b.forcing_rows, "plain") for c in range(K + W)]
        outs = out_buffers(torch, b.npad, Tc, b.ns, dev, "plain")
        flow_row = b.save_rows.index("PWATER_SURO" if op == "PERLND" else "IWATER_SURO")
        parts.append((op, sb, b, forc, outs, flow_row, torch.zeros(b.npad, dtype=torch.float64, device=dev)))
    torch.cuda.synchronize(dev)

    def step(c, accumulate):
        for _op, _sb, b, forc, outs, row, acc in parts:
            b.run_device(c * Tc, Tc, forc[c].data_ptr(), outs[c % 2].data_ptr(), stream.cuda_stream)
            if accumulate:
                with torch.cuda.stream(stream):
                    acc += outs[c % 2][:, row, :].sum(dim=0)

    for c in range(W):
        step(c, False)
    tdist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for c in range(W, W + K):
        step(c, True)
    e1.record(stream)
    tdist.barrier()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms = float(t.item())
    # the one collective of the job: per-segment runoff sums of every rank, in segment order, on every rank (NCCL all-gather)
    g0 = time.perf_counter()
    sums = [sb.gather(acc[: sb.n]) for _op, sb, _b, _f, _o, _r, acc in parts]
    torch.cuda.synchronize(dev)
    gather_ms = 1e3 * (time.perf_counter() - g0)
    res = {"workload": "one %d-segment ensemble = %d PERLND (SNOW+PWATER) + %d IMPLND (SNOW+IWATER), default SAVE sets, split with "
                       "dist.shard_bounds over %d ranks" % (n_total, half, half, world),
           "scaling": "strong", "segments_total": n_total,
           "window": "intervals %d..%d of the run (1976-01-17 .. 03-21: snow and melt season, the expensive regime; the headline "
                     "averages a whole year)" % (W * Tc, (W + K) * Tc), "segments_per_rank": [p[1].n for p in parts], "intervals_per_step": Tc,
           "steps": K, "warmup": W, "ms_per_step": ms / K, "value": n_total * Tc * K / (ms * 1e-3), "unit": "segment-timesteps/s",
           "kernels": [p[2].kernel_name() for p in parts],
           "bytes_per_segment_interval": {"PERLND": 240, "IMPLND": 176},
           "frac_of_measured_hbm_per_gpu": (half * 240 + half * 176) / world * Tc * K / (ms * 1e-3) / 1e9 / peak,
           "final_gather": {"what": "per-segment sums of SURO over the timed window, float64, all-gather over NCCL",
                            "elements": int(sum(x.numel() for x in sums)), "ms": gather_ms,
                            "checksum": float(sum(x.sum().item() for x in sums))}}
    for p in parts:
        p[1].close()
    del parts
    torch.cuda.empty_cache()
    return res


def run_watershed_stream(device, cal, n=5000, stations=16, years=10, chunk=24 * 61, jit=True):
    """BASELINE.json configs[4]: a testcbp-style watershed -- `n` full-chain PERLND land segments (the CBP segment's tables, option
    flags drawn per segment) on `stations` met gages with per-segment MFACTORs (shared-met mode: no per-segment forcing exists
    anywhere), `years` years hourly -- run END TO END through the product path: kernels in time chunks, the 10 series a CBP
    land segment hands its river model leave the device as float32 and are streamed chunk by chunk into the results store
    (hspsquared_b200.sink.NpyResults, a SupportsWriteTS / SupportsReadTS backend) while the next chunk computes.  Wall clock,
    everything included."""
    import shutil
    import tempfile

    import pandas as pd

    from hspsquared_b200 import synth
    from hspsquared_b200.engine import LandBatch
    from hspsquared_b200.sink import NpyResults, stream_batch

    steps = min(cal.steps, int(years * 8766))
    ens = synth.Ensemble("PERLND", n, seed=4321, base="cbp", options=FULLCHAIN_OPTIONS)
    forcings = list(synth.FULLCHAIN_FORCINGS)
    store = ens.store(cal)
    nf = len(forcings)
    s0 = synth.shared_series("PERLND", cal, 0, steps)
    series = np.empty((steps, nf, stations))
    for k in range(stations):  # gages differ by a temperature offset
        for r, nm in enumerate(forcings):
            a = s0["AIRTMP" if nm == "GATMP" else nm]
            if nm in ("AIRTMP", "GATMP"):
                a = a + (k - stations / 2) * 0.4
            elif nm == "DTMPG":
                a = (s0["AIRTMP"] + (k - stations / 2) * 0.4) - 5.0
            series[:, r, k] = a
    station = np.minimum((np.argsort(np.argsort(ens.u_airt)) * stations // n), stations - 1).astype(np.int32)
    mfac = np.ones((nf, n))
    mfac[forcings.index("PREC")] = 1.0 + 0.2 * ens.u_prec
    mfac[forcings.index("PETINP")] = 1.0 + 0.1 * ens.u_pet
    root = tempfile.mkdtemp(prefix="hsp2g_results_", dir="/dev/shm" if os.path.isdir("/dev/shm") and
                            shutil.disk_usage("/dev/shm").free > 3 * steps * n * 40 else None)
    try:
        tindex = pd.date_range(pd.Timestamp(str(cal.start)), periods=steps + 1, freq="60min")[1:]
        save_tables = {}
        for full in CBP10:
            act, name = full.split("_", 1)
            save_tables.setdefault(act, {})[name] = 1
        segs = ["P%05d" % (i + 1) for i in range(n)]
        t0 = time.perf_counter()
        sink = NpyResults(root, mode="w")
        with LandBatch(store, cal, forcings, list(CBP10), device=device, layout="plain", jit=jit, order="flags") as b:
            b.set_shared_forcing(series, station, mfac)
            t1 = time.perf_counter()
            stream_batch(sink, b, {"tindex": tindex}, "PERLND", segs, save_tables, None, chunk)
            kname = b.kernel_name()
            launches, _ = b.kernel_stats()
            errs = {k: int((v != 0).sum()) for k, v in b.errors().items() if (v != 0).any()}
        sink.close()
        t2 = time.perf_counter()
        back = NpyResults(root).read_ts("RESULT", "PERLND", segs[n // 2], "PWATER")
        nbytes = sum(os.path.getsize(os.path.join(root, f)) for f in os.listdir(root))
        return {"workload": "%d full-chain PERLND segments (CBP land segment, option flags drawn per segment) on %d met gages x MFACTOR, "
                            "%d hourly intervals (%.1f years), SAVE = the 10 CBP river-model series, streamed to the NpyResults store"
                            % (n, stations, steps, steps / 8766.0),
                "segments": n, "intervals": steps, "intervals_per_chunk": chunk, "kernel": kname, "gpu_launches": int(launches),
                "wall_s": t2 - t0, "setup_s": t1 - t0, "value": n * steps / (t2 - t1), "unit": "segment-timesteps/s",
                "stored_bytes": nbytes, "store_write_GBps": nbytes / (t2 - t1) / 1e9, "store": root.split("/")[1],
                "read_back": {"frame": list(back.shape), "columns": list(back.columns), "dtype": str(back.to_numpy().dtype)},
                "segments_with_error_counts": errs}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def run_extras(torch, dev, local_rank, rank, world, dist, cal, stream, peak):
    """The other BASELINE.json configurations, each a short device-resident measurement on this rank's GPU, so that the
    driver's record carries them beside the headline (rank 0 only)."""
    if rank != 0:
        return None
    out = {}
    for key, name, kw in (("implnd_100k", "implnd", {"order": "climate"}),
                          ("fullchain_250k_all84", "fullchain_all", {"K": 12, "order": "climate"}),
                          ("fullchain_250k_all84_tiled", "fullchain_all", {"K": 12, "order": "climate", "layout": "tiled"}),
                          ("fullchain_250k_cbp10", "fullchain_cbp10", {"K": 10, "order": "climate"}),
                          ("fullchain_250k_one_word_all84", "fullchain_uniform_all", {"K": 12, "order": "climate"}),
                          ("fullchain_250k_one_word_cbp10", "fullchain_uniform_cbp10", {"K": 10, "order": "climate"})):
        try:
            out[key] = run_config(torch, dev, name, cal, stream, peak, **kw)
        except Exception as ex:
            out[key] = {"error": repr(ex)}
    # configs[3] over one whole simulated year (the headline's measure; the entries above time the winter window)
    try:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
        import bench_config_year

        for key, layout in (("fullchain_250k_all84_year", "plain"), ("fullchain_250k_all84_year_tiled", "tiled")):
            try:
                out[key] = bench_config_year.one(torch, dev, "fullchain_all", layout, 128, 8784)
            except Exception as ex:
                out[key] = {"error": repr(ex)}
            torch.cuda.empty_cache()
    except Exception as ex:
        out["fullchain_250k_all84_year"] = {"error": repr(ex)}
    try:
        out["watershed_5k_10y_streamed"] = run_watershed_stream(local_rank, cal)
    except Exception as ex:
        out["watershed_5k_10y_streamed"] = {"error": repr(ex)}
    return out


if __name__ == "__main__":
    main()
