#!/bin/bash
OUT=gpurun_out/${1:-pow35}; mkdir -p $OUT
tools/_p35_probe | tee $OUT/probe.txt
timeout 600 python -m pytest tests/test_gpu_gen.py -m gpu -x -q -k "ldw or line_driven" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest.log | cut -c1-200
for G in 0 1 0 1; do
  if [ $G = 1 ]; then export PB200_LDW_GENERIC_POW=1; else unset PB200_LDW_GENERIC_POW; fi
  timeout 200 python bench.py --workload ldw --steps 40 --warmup 5 --no-e2e --no-cpu --no-secondary > $OUT/g$G.json 2> $OUT/g$G.err || tail -3 $OUT/g$G.err
  python -c "
import json; d=json.load(open('$OUT/g$G.json')); print('generic_pow=$G', round(d['ms_per_step'],4), d['roofline'].get('frac'))"
done
