#!/bin/bash
# reduced round-end validation (GPU budget): the scheduler tests, a whole-chunk-task run under overlap, the full bench line
O=gpurun_out/${1:-final_short}
mkdir -p $O
( timeout 90 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-ordered --no-alt-layout --no-extras --sub 438 ) > $O/bench_sub438.json 2> $O/bench_sub438.err; echo "sub438 rc=$?"
( timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "overlapped or subchunk or many_segments or threads" ) > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)"
( time timeout 420 python bench.py --steps 20 --warmup 5 ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; grep real $O/bench.err
