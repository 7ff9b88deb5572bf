#!/bin/bash
# A/B timing of fused-sweep tile shapes on the C4 bench (line-driven wind 1024 x 512) with the experiment library
# (PB200_EXTRA_DEFS=-DPB_GEN_TILES_EXP): tools/gpu_ab_ldw.sh TAG T0:T1 ...   (0 = the built-in default)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PB200_LIB=${AB_LIB:-$PWD/pluto_sirocco_b200/lib/libplutob200_gx.so}
for V in "$@"; do
  export PB200_GEN_TILE0=${V%%:*} PB200_GEN_TILE1=${V#*:}
  timeout 200 python bench.py --workload ldw --steps 40 --warmup 5 --no-e2e --no-cpu --no-secondary > $OUT/$V.json 2> $OUT/$V.err || tail -3 $OUT/$V.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$V.json")); print("$V", "C4", round(d["ms_per_step"],4), d.get("gpu_launches"), d["roofline"].get("frac"))
except Exception as e: print("$V", "failed", e)
PY
done
if [ -n "$AB_TEST" ]; then
  timeout 600 python -m pytest tests/test_gpu_gen.py -m gpu -x -q -k "$AB_TEST" > $OUT/pytest.log 2>&1; tail -n 2 $OUT/pytest.log
fi
