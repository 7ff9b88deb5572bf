#!/bin/bash
# A/B timing of variant libraries on the C2 bench (and a parity spot check): tools/gpu_ab.sh TAG lib1 lib2 ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for L in "$@"; do
  export PB200_LIB=$PWD/pluto_sirocco_b200/lib/$L
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/$L.json 2> $OUT/$L.err || tail -3 $OUT/$L.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$L.json")); print("$L", "C2", round(d["ms_per_step"],3), d["roofline"]["step"]["frac"], {k:round(v,3) for k,v in d["roofline"]["kernels_ms"].items()})
except Exception as e: print("$L", "failed", e)
PY
  if [ -n "$AB_C5" ]; then
  timeout 300 python bench.py --size 256 --recon PARABOLIC --rk RK3 --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/$L.c5.json 2> $OUT/$L.c5.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$L.c5.json")); print("$L", "C5", round(d["ms_per_step"],3), d["roofline"]["step"]["frac"])
except Exception as e: print("$L", "failed", e)
PY
  fi
  if [ -n "$AB_TEST" ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$AB_TEST" > $OUT/$L.pytest.log 2>&1; tail -n 2 $OUT/$L.pytest.log
  fi
done
