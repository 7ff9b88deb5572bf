#!/bin/bash
# Quick GPU visit: parity tests (optional -k filter in $2) and a short bench.
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$2" > $OUT/pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
fi
echo "pytest exit $?" >> $OUT/pytest.log
tail -30 $OUT/pytest.log
if [ "$3" != "nobench" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench.json 2> $OUT/bench.err
  echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
fi
