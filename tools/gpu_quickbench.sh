#!/bin/bash
# fast-path parity tests + C2 bench (+ LDW bench, tests)
OUT=gpurun_out/${1:-qb}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest.log 2>&1; tail -n 3 $OUT/pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json")); print("C2", round(d["ms_per_step"],3), round(d["value"],1), d["roofline"]["step"]["frac"], {k:round(v,3) for k,v in d["roofline"]["kernels_ms"].items()})
PY
timeout 300 python bench.py --size 256 --recon PARABOLIC --rk RK3 --steps 20 --warmup 3 --no-e2e --no-cpu > $OUT/c5.json 2> $OUT/c5.err
timeout 300 python bench.py --workload ldw --steps 50 --warmup 5 > $OUT/ldw.json 2> $OUT/ldw.err
python - <<PY
import json
for f in ("c5","ldw"):
    d=json.load(open("$OUT/%s.json"%f)); print(f, round(d["ms_per_step"],3), round(d["value"],1))
PY
timeout 600 python -m pytest tests/test_gpu_gen.py -m gpu -x -q -k "ldw or cooling" > $OUT/pytest2.log 2>&1; tail -n 2 $OUT/pytest2.log
