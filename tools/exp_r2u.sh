#!/bin/bash
# lock-step blocks (jit.py geometry: HSP2G_JIT_BLOCK threads per block, HSP2G_JIT_LOCKSTEP barrier level) on the plain headline,
# one simulated year; first the bit-exactness of the lock-step scheduler (sub-chunk tasks, padding, mixed option words)
mkdir -p gpurun_out/r2u
HSP2G_JIT_BLOCK=512 HSP2G_JIT_LOCKSTEP=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "subchunk or many_segments or categorical or water_balance" > gpurun_out/r2u/pytest_lockstep.log 2>&1
echo "pytest lockstep rc=$?"
tail -3 gpurun_out/r2u/pytest_lockstep.log
for v in "64 0" "512 1" "256 1" "128 1" "512 2" "256 2" "512 0" "256 0"; do
  set -- $v
  HSP2G_JIT_BLOCK=$1 HSP2G_JIT_LOCKSTEP=$2 timeout 200 python tools/exp_config.py headline --order climate --layout plain \
    > gpurun_out/r2u/b$1_l$2.jsonl 2> gpurun_out/r2u/b$1_l$2.err
  echo "$v rc=$? $(python -c "import json,sys; r=json.loads(open('gpurun_out/r2u/b$1_l$2.jsonl').readline()); print(r.get('ms_per_step'), r.get('frac_of_measured_hbm'), r.get('error'))" 2>&1)"
done
