#!/usr/bin/env python
"""Copy-only PCIe microbenchmark: what host<->device bandwidth do R concurrent ranks (one process per GPU) get on this box?

  python tools/pcie_bench.py --ranks 1,2,4,8 --out profiles/pcie_ranks.json

For every rank count R the parent starts R worker processes (GPU r each, all released by a file barrier) and every worker
times, with CUDA events on its own stream, H2D alone, D2H alone and both directions at once, for each kind of host buffer:
  default     cudaHostAlloc(cudaHostAllocDefault)          (hsp2g_host_alloc)
  wc          cudaHostAlloc(cudaHostAllocWriteCombined)    (H2D source only; D2H target stays default)
  registered  malloc + first touch + cudaHostRegister       (a numpy array pinned in place)
  pageable    plain malloc'd memory                          (cudaMemcpyAsync stages it through the driver)
optionally with the worker pinned (sched_setaffinity) to the cores of the GPU's NUMA node before it allocates.  The per-rank
figures and their sum per R say whether the end-to-end scaling of bench.py is bounded by the platform (host memory / IOMMU /
PCIe switch topology) or by the library.  Uses torch only for the CUDA runtime (pinned allocation flags come from cudart via
the C ABI of libhsp2g).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_numa_cpus(index):
    import bench

    return bench.gpu_local_cpus(index)


def worker(rank, nranks, nbytes, reps, barrier_dir, affinity):
    import numpy as np
    import torch

    from hspsquared_b200 import _lib

    cpus = None
    if affinity:
        cpus = gpu_numa_cpus(rank)
        if cpus:
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
            if allowed:
                os.sched_setaffinity(0, allowed)
                cpus = allowed
    torch.cuda.set_device(rank)
    lib = _lib.load()
    dev = torch.device("cuda", rank)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def alloc(kind):
        if kind in ("default", "wc"):
            p = lib.hsp2g_host_alloc_flags(nbytes, _lib.HOST_WRITE_COMBINED if kind == "wc" else 0)
            if not p:
                raise MemoryError(lib.hsp2g_last_error().decode())
            return p, lambda: lib.hsp2g_host_free(p)
        a = np.empty(nbytes, dtype=np.uint8)
        a[::4096] = 1  # first touch on this rank's cores
        if kind == "registered":
            _lib.check(lib.hsp2g_host_register(a.ctypes.data, nbytes))
            return a.ctypes.data, lambda a=a: lib.hsp2g_host_unregister(a.ctypes.data)
        return a.ctypes.data, lambda a=a: None

    # copies are issued through the CUDA runtime library itself
    rtl = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rtl = C.CDLL(name)
            break
        except OSError:
            continue
    if rtl is None:
        import glob

        for path in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + \
                glob.glob("/usr/local/cuda/lib64/libcudart.so*"):
            try:
                rtl = C.CDLL(path)
                break
            except OSError:
                continue
    rtl.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]

    def memcpy(dst, src, kind, stream):
        e = rtl.cudaMemcpyAsync(dst, src, nbytes, kind, stream.cuda_stream)
        if e:
            raise RuntimeError("cudaMemcpyAsync -> %d" % e)

    def wait_all(tag):
        open(os.path.join(barrier_dir, "%s.%d" % (tag, rank)), "w").close()
        t0 = time.time()
        while sum(1 for f in os.listdir(barrier_dir) if f.startswith(tag + ".")) < nranks:
            time.sleep(0.0005)
            if time.time() - t0 > 120:
                raise TimeoutError("barrier " + tag)

    res = {"rank": rank, "cpus": (len(cpus) if cpus else None)}
    for kind in (("default", "wc", "registered", "pageable") if nranks == 1 else ("default", "wc")):
        h_src, free_src = alloc(kind)
        h_dst, free_dst = alloc("default" if kind == "wc" else kind)
        r = {}
        for mode in ("h2d", "d2h", "bidir"):
            n = reps if kind != "pageable" else max(1, reps // 3)
            for it in range(2):  # one warm-up pass, one timed
                wait_all("%s_%s_%d" % (kind, mode, it))
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for _ in range(n if it else 1):
                    if mode in ("h2d", "bidir"):
                        memcpy(d_in.data_ptr(), h_src, 1, s_in)
                    if mode in ("d2h", "bidir"):
                        memcpy(h_dst, d_out.data_ptr(), 2, s_out)
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
            r[mode] = round(nbytes * n * (2 if mode == "bidir" else 1) / dt / 1e9, 2)
        res[kind] = r
        free_src()
        free_dst()
    print("RESULT " + json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", default="1")
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--out", default=None)
    ap.add_argument("--worker", type=int, default=None)
    ap.add_argument("--nranks", type=int, default=1)
    ap.add_argument("--barrier", default=None)
    ap.add_argument("--affinity", type=int, default=0)
    a = ap.parse_args()
    if a.worker is not None:
        worker(a.worker, a.nranks, a.mb << 20, a.reps, a.barrier, bool(a.affinity))
        return
    import tempfile

    import torch

    ngpu = torch.cuda.device_count()
    report = {"gpus_visible": ngpu, "host_cpus": len(os.sched_getaffinity(0)), "mb_per_copy": a.mb, "reps": a.reps,
              "numa_cpus_of_gpu": {str(i): (len(gpu_numa_cpus(i)) if gpu_numa_cpus(i) else None) for i in range(ngpu)},
              "unit": "GB/s per rank (bidir = sum of both directions)", "runs": []}
    for R in [int(x) for x in a.ranks.split(",")]:
        if R > ngpu:
            continue
        for aff in (0, 1):
            d = tempfile.mkdtemp(prefix="pcie_barrier_")
            procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(r), "--nranks", str(R), "--mb",
                                       str(a.mb), "--reps", str(a.reps), "--barrier", d, "--affinity", str(aff)],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(R)]
            ranks = []
            for p in procs:
                out, err = p.communicate(timeout=600)
                got = [json.loads(l[7:]) for l in out.splitlines() if l.startswith("RESULT ")]
                ranks.append(got[0] if got else {"error": err[-500:]})
            run = {"ranks": R, "affinity": bool(aff), "per_rank": ranks}
            ok = [r for r in ranks if "error" not in r]
            if ok:
                run["sum"] = {k: {m: round(sum(r[k][m] for r in ok), 1) for m in ("h2d", "d2h", "bidir")}
                              for k in ("default", "wc", "registered", "pageable") if all(k in r for r in ok)}
            report["runs"].append(run)
            print(json.dumps({"ranks": R, "affinity": bool(aff), "sum": run.get("sum")}), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
