#!/usr/bin/env python
"""One named BASELINE configuration, device-resident, with the knobs of bench.run_config on the command line (used for A/B
experiments and for ncu captures, which want few launches):

  python tools/exp_config.py fullchain_all --layout tiled --order climate --steps 6 --warmup 3
  ncu --set full -k regex:'hsp2g_jit_kernel|perlnd_kernel' -s 3 -c 1 ... python tools/exp_config.py headline --steps 2
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--layout", default="plain")
    ap.add_argument("--order", default="flags")
    ap.add_argument("--no-jit", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=None)
    ap.add_argument("--sub", type=int, default=None)
    ap.add_argument("--segments", type=int, default=None)
    ap.add_argument("--first-chunk", type=int, default=0)
    ap.add_argument("--f32out", action="store_true")
    ap.add_argument("--f32in", action="store_true")
    a = ap.parse_args()
    import numpy as np
    import torch

    from hspsquared_b200.calendar import Calendar

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cal = Calendar(B.RUN_START, B.RUN_STOP, 60, "ns")
    stream = torch.cuda.Stream(dev)
    peak = B.measured_peak()[0]
    for name in a.configs:
        for layout in a.layout.split(","):
            for order in a.order.split(","):
                try:
                    r = B.run_config(torch, dev, name, cal, stream, peak, n=a.segments, layout=layout, jit=not a.no_jit,
                                     order=None if order == "none" else order, K=a.steps, W=a.warmup, Tc=a.chunk, sub=a.sub,
                                     out_dtype=np.float32 if a.f32out else None, first_chunk=a.first_chunk,
                                     forcing_dtype=np.float32 if a.f32in else None)
                except Exception as ex:
                    r = {"config": name, "layout": layout, "order": order, "error": repr(ex)}
                print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
