#!/bin/bash
# full GPU test suite + the default bench (both arms): tools/gpu_full.sh TAG
OUT=gpurun_out/${1:-full}; mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=10 ) > $OUT/pytest.log 2>&1
echo "pytest exit $?"; tail -18 $OUT/pytest.log
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; tail -2 $OUT/bench.err; cat $OUT/bench.json
( time timeout 900 python bench.py --impl reference ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "ref exit $?"; tail -3 $OUT/bench_ref.err; cat $OUT/bench_ref.json
