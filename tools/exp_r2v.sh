#!/bin/bash
# how much of a headline launch is its tail?  same year, launches of 438 / 876 / 1752 intervals, and task lengths around 64
mkdir -p gpurun_out/r2v
run() { # name, args...
  n=$1; shift
  timeout 300 python tools/exp_config.py headline --order climate --layout plain "$@" > gpurun_out/r2v/$n.jsonl 2> gpurun_out/r2v/$n.err
  echo "$n rc=$? $(python -c "import json,sys; r=json.loads(open('gpurun_out/r2v/$n.jsonl').readline()); print(r.get('intervals_per_step'), r.get('ms_per_step'), r.get('frac_of_measured_hbm'), r.get('error'))" 2>&1)"
}
run c438 --steps 20 --warmup 3
run c876 --chunk 876 --steps 10 --warmup 2
run c1752 --chunk 1752 --steps 5 --warmup 1
run c438_s73 --steps 20 --warmup 3 --sub 73
run c438_s55 --steps 20 --warmup 3 --sub 55
run c438_s44 --steps 20 --warmup 3 --sub 44
run c438_s110 --steps 20 --warmup 3 --sub 110
run c876_s73 --chunk 876 --steps 10 --warmup 2 --sub 73
