#!/usr/bin/env python
"""Why do the per-option-word kernels of a grouped ensemble not overlap?  Times (a) every group's kernel alone, (b) all groups
through hsp2g_run_device_group, (c) all groups launched from here on torch streams."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    import numpy as np
    import torch

    from hspsquared_b200.calendar import Calendar
    from hspsquared_b200.engine import GroupedLandBatch

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cal = Calendar(B.RUN_START, B.RUN_STOP, 60, "ns")
    name = sys.argv[1] if len(sys.argv) > 1 else "fullchain_all"
    ens, forcings, save, n, Tc, _ = B.config_workload(name)
    store = ens.store(cal)
    g = GroupedLandBatch(store, cal, forcings, save, device=0, layout="plain", jit=True)
    W = g.width
    parts = [(c0, b.npad, ens.take(idx)) for c0, idx, b in zip(g.first_column, g.members, g.groups)]
    nchunks = 5
    forc = []
    for c in range(nchunks):
        big = torch.empty((Tc, g.nf, W), dtype=torch.float64, device=dev)
        for c0, w_, e_ in parts:
            big[:, :, c0:c0 + w_] = B.device_forcing(torch, e_, cal, c * Tc, Tc, w_, dev, g.forcing_rows, "plain")
        forc.append(big)
    out = torch.empty((Tc, g.ns, W), dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # warm-up + (b) grouped
    res = {"groups": len(g.groups), "sizes": [int(len(m)) for m in g.members][:6]}
    for c in range(2):
        g.run_device(c * Tc, Tc, forc[c].data_ptr(), out.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(st)
    g.run_device(2 * Tc, Tc, forc[2].data_ptr(), out.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    res["grouped_ms"] = e0.elapsed_time(e1)
    # (a) each group alone on the same stream, one after the other
    alone = []
    for k, (b, c0) in enumerate(zip(g.groups, g.first_column)):
        e0, e1 = ev(), ev()
        e0.record(st)
        b.run_device(3 * Tc, Tc, forc[3].data_ptr() + c0 * 8, out.data_ptr() + c0 * 8, st.cuda_stream, ldf=W, ldo=W)
        e1.record(st)
        torch.cuda.synchronize()
        alone.append(e0.elapsed_time(e1))
    res["alone_ms_sum"] = sum(alone)
    res["alone_ms"] = [round(x, 3) for x in alone[:8]]
    # (c) from here: one torch stream per group, no library fork / join
    streams = [torch.cuda.Stream(dev) for _ in g.groups]
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(st)
    for s_ in streams:
        s_.wait_stream(st)
    for b, c0, s_ in zip(g.groups, g.first_column, streams):
        b.run_device(4 * Tc, Tc, forc[4].data_ptr() + c0 * 8, out.data_ptr() + c0 * 8, s_.cuda_stream, ldf=W, ldo=W)
    for s_ in streams:
        st.wait_stream(s_)
    e1.record(st)
    torch.cuda.synchronize()
    res["torch_streams_ms"] = e0.elapsed_time(e1)
    res["max_connections"] = os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
