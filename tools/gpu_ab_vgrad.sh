#!/bin/bash
# gen_vgrad loop variants on C4: the shipped library first, then the experiment library (-DPB_VGRAD_EXP) with
# PB200_VGRAD_VAR = 10 * bins per iteration + resident blocks the registers are cut for; LDW parity tests for $AB_TESTS
OUT=gpurun_out/${1:-abvgrad}; mkdir -p $OUT; shift
run() { timeout 100 python bench.py --workload ldw --steps 40 --warmup 5 --no-e2e --no-cpu --no-secondary > $OUT/$1.json 2> $OUT/$1.err
  python -c "
import json; d=json.load(open('$OUT/$1.json')); print('$1', round(d['ms_per_step'],4))" 2>/dev/null || echo "$1 failed"; }
run shipped
export PB200_LIB=$PWD/pluto_sirocco_b200/lib/libplutob200_gx.so
for V in "$@"; do PB200_VGRAD_VAR=$V run v$V; done
for V in $AB_TESTS; do
  PB200_VGRAD_VAR=$V timeout 120 python -m pytest tests/test_gpu_gen.py -m gpu -x -q -k "ldw or line_driven" > $OUT/pytest_$V.log 2>&1; echo "tests $V: $(tail -n 1 $OUT/pytest_$V.log)"
done
