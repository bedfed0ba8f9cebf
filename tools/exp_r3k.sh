#!/bin/bash
# headline with overlapping launches: intervals per (tile, sub-chunk) task
mkdir -p gpurun_out/r3k
for sub in ${SUBS:-64 73 110 146 219}; do
  timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-ordered --no-alt-layout --no-extras --sub $sub > gpurun_out/r3k/sub$sub.json 2> gpurun_out/r3k/sub$sub.err
  echo "sub $sub rc=$? $(python -c "
import json
d=json.loads(open('gpurun_out/r3k/sub$sub.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity']['bit_identical'])
" 2>&1)"
done
