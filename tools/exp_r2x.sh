#!/bin/bash
# SNOW: per-segment constant heads hoisted to the host, division by 24 / 100 / 144 in three operations
mkdir -p gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_bit_exact or golden_reference or float32_outputs or overlapped" > gpurun_out/r2x/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2x/pytest.log
run() { # name, args...
  n=$1; shift
  timeout 300 python tools/exp_config.py "$@" > gpurun_out/r2x/$n.jsonl 2> gpurun_out/r2x/$n.err
  echo "$n rc=$? $(python -c "
import json
for l in open('gpurun_out/r2x/$n.jsonl'):
    r=json.loads(l); print(r.get('config'), r.get('series_layout'), r.get('intervals_per_step'), round(r.get('ms_per_step',0),4), round(r.get('frac_of_measured_hbm',0),4), r.get('error'))
" 2>&1 | tr '\n' ';')"
}
run head headline --order climate --layout plain,tiled
run head2 headline --order climate --layout plain
run impl implnd --order climate --layout plain
run chain fullchain_all fullchain_cbp10 --order climate --layout plain --steps 12
