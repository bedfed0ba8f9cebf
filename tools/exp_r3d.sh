#!/bin/bash
# full chain, all 84 outputs (BASELINE configs[3]): plain against tiled series layout, block sizes (adjacent tiles per block write
# adjacent 256-byte pieces of a row at about the same time), store cache policy
mkdir -p gpurun_out/r3d
run() { # name, env..., -- args
  n=$1; shift
  env "${@:1:$(($#-0))}" true 2>/dev/null
}
go() { # name "ENV=.. ENV=.." args...
  n=$1; e=$2; shift 2
  env $e timeout 300 python tools/exp_config.py "$@" > gpurun_out/r3d/$n.jsonl 2> gpurun_out/r3d/$n.err
  echo "$n rc=$? $(python -c "
import json
for l in open('gpurun_out/r3d/$n.jsonl'):
    r=json.loads(l); print(r.get('config'), r.get('series_layout'), r.get('registers'), round(r.get('ms_per_step',0),4), round(r.get('frac_of_measured_hbm',0),4), r.get('error'))
" 2>&1 | tr '\n' ';')"
}
go base "X=1" fullchain_all fullchain_uniform_all --order climate --layout plain,tiled --steps 12
go b128 "HSP2G_JIT_BLOCK=128" fullchain_all fullchain_uniform_all --order climate --layout plain --steps 12
go b256 "HSP2G_JIT_BLOCK=256" fullchain_all fullchain_uniform_all --order climate --layout plain --steps 12
go pol1 "HSP2G_JIT_FLAGS=-DHSP_STORE_POLICY=1" fullchain_all fullchain_uniform_all --order climate --layout plain --steps 12
go pol2 "HSP2G_JIT_FLAGS=-DHSP_STORE_POLICY=2" fullchain_all --order climate --layout plain --steps 12
