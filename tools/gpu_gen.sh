#!/bin/bash
# general-path parity tests + the C4 bench: tools/gpu_gen.sh TAG
OUT=gpurun_out/${1:-gen}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gen.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest.log | cut -c1-200
timeout 300 python bench.py --workload ldw --steps 40 --warmup 5 --no-cpu --no-secondary > $OUT/bench_ldw.json 2> $OUT/bench_ldw.err || tail -3 $OUT/bench_ldw.err
python -c "
import json; d=json.load(open('$OUT/bench_ldw.json')); print('C4', round(d['ms_per_step'],4), d['value'], d['roofline'].get('frac'), d.get('e2e',{}).get('value'))"
