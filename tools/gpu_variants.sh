#!/bin/bash
# bench every kernel-variant library in pluto_sirocco_b200/lib (perf experiments)
OUT=gpurun_out/variants
mkdir -p $OUT
for lib in pluto_sirocco_b200/lib/libplutob200.so pluto_sirocco_b200/lib/libv_*.so; do
  name=$(basename $lib .so)
  PB200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json")); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernels_ms"].items()})
except Exception as e: print("$name failed", e)
PY
done
