#!/bin/bash
# launch overlap (hsp2g_set_launch_overlap): bit-exactness test, then the headline and the other configurations with and without
mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "overlapped or subchunk or many_segments" > gpurun_out/r2w/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2w/pytest.log
run() { # name, overlap, args...
  n=$1; ov=$2; shift; shift
  HSP2G_OVERLAP=$ov timeout 300 python tools/exp_config.py "$@" > gpurun_out/r2w/$n.jsonl 2> gpurun_out/r2w/$n.err
  echo "$n rc=$? $(python -c "
import json
for l in open('gpurun_out/r2w/$n.jsonl'):
    r=json.loads(l); print(r.get('config'), r.get('series_layout'), r.get('intervals_per_step'), round(r.get('ms_per_step',0),4), round(r.get('frac_of_measured_hbm',0),4), r.get('error'))
" 2>&1 | tr '\n' ';')"
}
run head_ov1 1 headline --order climate --layout plain,tiled
run head_ov0 0 headline --order climate --layout plain,tiled
run head_ov1_caller 1 headline --order none --layout plain
run head_ov0_caller 0 headline --order none --layout plain
run impl_ov1 1 implnd --order climate --layout plain
run impl_ov0 0 implnd --order climate --layout plain
run chain_ov1 1 fullchain_all fullchain_cbp10 --order climate --layout plain --steps 12
run chain_ov0 0 fullchain_all fullchain_cbp10 --order climate --layout plain --steps 12
