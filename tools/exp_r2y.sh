#!/bin/bash
# 18 resident warps per SM?  float32 forcings halve the ring (48 rows of shared memory per warp instead of 54), which leaves room
# for 9 two-warp blocks if the kernel fits 112 registers (44 B of spill): measured against the shipped 128 registers / 16 warps
mkdir -p gpurun_out/r2y
run() { # name, args...
  n=$1; shift
  timeout 300 python tools/exp_config.py "$@" > gpurun_out/r2y/$n.jsonl 2> gpurun_out/r2y/$n.err
  echo "$n rc=$? $(python -c "
import json
for l in open('gpurun_out/r2y/$n.jsonl'):
    r=json.loads(l); print(r.get('config'), r.get('series_layout'), r.get('registers'), round(r.get('ms_per_step',0),4), round(r.get('frac_of_measured_hbm',0),4), r.get('kernel','')[-40:], r.get('error'))
" 2>&1 | tr '\n' ';')"
}
run f32_128 headline --order climate --layout plain --f32in
HSP2G_JIT_MAXNREG=1 HSP2G_JIT_MAXREG=112 run f32_112 headline --order climate --layout plain --f32in
HSP2G_JIT_MAXNREG=1 HSP2G_JIT_MAXREG=120 run f32_120 headline --order climate --layout plain --f32in
HSP2G_JIT_MAXNREG=1 HSP2G_JIT_MAXREG=128 run f32_128x headline --order climate --layout plain --f32in
run chain fullchain_all fullchain_cbp10 --order climate --layout plain --steps 12
run chain2 fullchain_all fullchain_cbp10 --order climate --layout plain --steps 12
