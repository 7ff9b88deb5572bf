#!/bin/bash
# A/B timing of general-path kernel variants on the C4 bench (line-driven wind 1024 x 512) with an experiment library
# that selects them by environment: tools/gpu_ab_ldw.sh TAG "VAR=val VAR2=val" ...   ("-" = defaults)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PB200_LIB=${AB_LIB:-$PWD/pluto_sirocco_b200/lib/libplutob200_gx.so}
n=0
for V in "$@"; do
  n=$((n+1))
  if [ "$V" = "-" ]; then E=""; else E="$V"; fi
  env $E timeout 200 python bench.py --workload ldw --steps 40 --warmup 5 --no-e2e --no-cpu --no-secondary > $OUT/v$n.json 2> $OUT/v$n.err || tail -3 $OUT/v$n.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/v$n.json")); print("$V", "C4", round(d["ms_per_step"],4), d.get("gpu_launches"), d["roofline"].get("frac"))
except Exception as e: print("$V", "failed", e)
PY
done
if [ -n "$AB_TEST" ]; then
  for V in "$@"; do
    if [ "$V" = "-" ]; then E=""; else E="$V"; fi
    env $E timeout 600 python -m pytest tests/test_gpu_gen.py -m gpu -x -q -k "$AB_TEST" > $OUT/pytest.log 2>&1; echo "$V: $(tail -n 1 $OUT/pytest.log)"
  done
fi
