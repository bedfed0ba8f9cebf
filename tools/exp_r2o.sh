#!/bin/bash
# register cap of the full-chain kernels (BASELINE configs[3]): default (168, raised on spills) vs 144 vs 128
mkdir -p gpurun_out/r2o
for mr in default 144 128; do
  if [ $mr = default ]; then unset HSP2G_JIT_MAXREG; else export HSP2G_JIT_MAXREG=$mr; fi
  timeout 400 python tools/exp_config.py fullchain_all fullchain_cbp10 fullchain_uniform_all fullchain_uniform_cbp10 --order climate --steps 8 --warmup 3 \
     > gpurun_out/r2o/chain_$mr.jsonl 2> gpurun_out/r2o/chain_$mr.err
  echo "$mr rc=$?"
done
