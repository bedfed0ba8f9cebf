#!/bin/bash
# IMPLND SNOW+IWATER (106 registers at the 128 cap: 9 blocks per SM): lower register caps = more resident warps
mkdir -p gpurun_out/r3i
for cap in 128 96 80; do
  HSP2G_JIT_MAXREG=$cap timeout 200 python tools/exp_config.py implnd --order climate --layout plain > gpurun_out/r3i/implnd_$cap.jsonl 2> gpurun_out/r3i/implnd_$cap.err
  echo "cap $cap rc=$? $(python -c "
import json
for l in open('gpurun_out/r3i/implnd_$cap.jsonl'):
    r=json.loads(l); print(r.get('registers'), round(r.get('ms_per_step',0),4), round(r.get('frac_of_measured_hbm',0),4), r.get('error'))
" 2>&1 | tr '\n' ';')"
done
