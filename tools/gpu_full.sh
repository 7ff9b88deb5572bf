#!/bin/bash
# Full GPU visit: whole -m gpu suite, C2 bench, LDW bench + launch list
OUT=gpurun_out/${1:-full}
mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -x -q -rs > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; cut -c1-600 $OUT/bench.json
timeout 600 python bench.py --workload ldw --steps 50 --warmup 5 > $OUT/bench_ldw.json 2> $OUT/bench_ldw.err; cut -c1-330 $OUT/bench_ldw.json; tail -2 $OUT/bench_ldw.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_ldw.csv \
   python bench.py --workload ldw --steps 3 --warmup 3 > $OUT/launches_ldw.log 2>&1
