#!/bin/bash
# task length (intervals per (tile, sub-chunk) task) on the plain headline, one simulated year
mkdir -p gpurun_out/r2q
for sub in 64 32 73 110 146 219 438; do
  timeout 200 python tools/exp_config.py headline --order climate --layout plain,tiled --sub $sub > gpurun_out/r2q/sub_$sub.jsonl 2> gpurun_out/r2q/sub_$sub.err
  echo "$sub rc=$?"
done
