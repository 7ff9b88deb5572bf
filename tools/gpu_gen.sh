#!/bin/bash
# general path: parity tests + LDW bench, fused and unfused: tools/gpu_gen.sh TAG
OUT=gpurun_out/${1:-gen}; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_gen.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest.log
for F in 0 1; do
PB200_GEN_FUSED=$F timeout 300 python bench.py --workload ldw --steps 50 --warmup 5 > $OUT/ldw_f$F.json 2> $OUT/ldw_f$F.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/ldw_f$F.json")); print("fused=$F ldw", round(d["ms_per_step"],4), "ms", round(d["value"],1), "Mz/s frac", round(d["roofline"]["frac"],4), "launches/step", d["gpu_launches"]/d["steps"])
except Exception as e: print("ldw failed", e); print(open("$OUT/ldw_f$F.err").read()[-1500:])
PY
done
