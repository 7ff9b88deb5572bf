#!/bin/bash
# C4 / C5 / drop-in visit: LDW drop-in test, LDW bench (+ launch list), C5 bench.
OUT=gpurun_out/c4
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "dropin_line_driven" -rs > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
timeout 600 python bench.py --workload ldw --steps 50 --warmup 5 > $OUT/bench_ldw.json 2> $OUT/bench_ldw.err
cat $OUT/bench_ldw.json; tail -3 $OUT/bench_ldw.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_ldw.csv \
   python bench.py --workload ldw --steps 3 --warmup 3 > $OUT/launches_ldw.log 2>&1
timeout 600 python bench.py --size 256 --recon PARABOLIC --rk RK3 --steps 20 --warmup 3 --no-cpu > $OUT/bench_c5.json 2> $OUT/bench_c5.err
cat $OUT/bench_c5.json | cut -c1-400; tail -3 $OUT/bench_c5.err
