#!/bin/bash
OUT=gpurun_out/${1:-e2e}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_call" > $OUT/pytest.log 2>&1; tail -n 15 $OUT/pytest.log
for p in 16 24 32 48; do
PB200_HOST_PIPELINE=$p timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 4 > $OUT/bench_p$p.json 2> $OUT/bench_p$p.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_p$p.json")); print("host pipeline $p: e2e", round(d["e2e"]["value"],1), "value", round(d["value"],1))
except Exception as e: print("failed", e, open("$OUT/bench_p$p.err").read()[-600:])
PY
done
