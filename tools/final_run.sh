#!/bin/bash
# round-end validation on one B200: GPU test suite, smoke, both bench arms, launch list and one full capture of the headline kernel
O=gpurun_out/${1:-final}
mkdir -p $O
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)"; grep real $O/pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > $O/smoke.log 2>&1; echo "smoke rc=$? $(grep -c 'smoke ok' $O/smoke.log)"
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; grep real $O/bench.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; grep real $O/bench_ref.err
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-ordered --no-alt-layout --no-extras"
( timeout 300 $B --no-overlap ) > $O/bench_no_overlap.json 2> $O/bench_no_overlap.err; echo "bench --no-overlap rc=$?"
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $B ) > $O/ncu_l.log 2>&1; echo "launches rc=$?"
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:hsp2g_jit_kernel -s 12 -c 1 -o $O/ncu_step438 $B ) > $O/ncu_f.log 2>&1; echo "ncu rc=$?"
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:hsp2g_jit_kernel -s 3 -c 1 -o $O/ncu_march $B ) > $O/ncu_m.log 2>&1; echo "ncu march rc=$?"
