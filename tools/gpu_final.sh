#!/bin/bash
# end-of-round visit: full GPU suite, smoke, default bench (both arms), ncu launch list of the bench command
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -q --durations=8 ) > $OUT/pytest.log 2>&1
echo "pytest exit $?"; tail -16 $OUT/pytest.log | cut -c1-180
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log | cut -c1-200
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; tail -2 $OUT/bench.err; cut -c1-1500 $OUT/bench.json
( time timeout 900 python bench.py --impl reference ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "ref exit $?"; tail -3 $OUT/bench_ref.err; cut -c1-600 $OUT/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_default.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > $OUT/ncu_bench.log 2>&1; echo "ncu exit $?"
