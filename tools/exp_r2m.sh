#!/bin/bash
# register cap x division inlining on the headline kernel (bench.py, 20 chunks = one simulated year)
mkdir -p gpurun_out/r2m
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-extras --no-alt-layout --no-ordered"
for v in "128 1" "128 0" "96 0" "96 1" "80 0"; do
  set -- $v
  HSP2G_JIT_MAXREG=$1 HSP2G_JIT_INLINE_DIV=$2 timeout 200 $B > gpurun_out/r2m/bench_$1_$2.json 2> gpurun_out/r2m/bench_$1_$2.err
  echo "$v rc=$?"
done
