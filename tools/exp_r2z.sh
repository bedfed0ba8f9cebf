#!/bin/bash
# cache policy of the saved-series stores (HSP_STORE_POLICY: 0 = .cs streaming, shipped; 1 = write-back; 2 = .cg; 3 = .wt), both layouts
mkdir -p gpurun_out/r2z
for pol in 0 1 2 3; do
  HSP2G_JIT_FLAGS="-DHSP_STORE_POLICY=$pol" timeout 300 python tools/exp_config.py headline --order climate --layout plain,tiled \
    > gpurun_out/r2z/pol$pol.jsonl 2> gpurun_out/r2z/pol$pol.err
  echo "policy $pol rc=$? $(python -c "
import json
for l in open('gpurun_out/r2z/pol$pol.jsonl'):
    r=json.loads(l); print(r.get('series_layout'), round(r.get('ms_per_step',0),4), round(r.get('frac_of_measured_hbm',0),4), r.get('error'))
" 2>&1 | tr '\n' ';')"
done
