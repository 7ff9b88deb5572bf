#!/bin/bash
# the 8-GPU visit: multi-GPU correctness tests, the driver's scaling bench at N GPUs, the C-side multi bench
OUT=gpurun_out/${1:-scale8}; N=${2:-8}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_slab_nccl_gpu.py tests/test_dropin_gpu.py -m gpu -x -q -k "slab or multi" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n$N.json")); print("N=$N value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
    for k,v in (d.get("secondary") or {}).items(): print("   ", k, round(v.get("value",0),1), "ms", round(v.get("ms_per_step",0),3), "frac", v.get("roofline_step_frac"))
except Exception as e: print("bench failed", e); print(open("$OUT/bench_n$N.err").read()[-1500:])
PY
timeout 600 python tools/multi_bench.py --ngpus $N --size 512 --host-size 256 > $OUT/multi_c_n$N.json 2> $OUT/multi_c_n$N.err; cat $OUT/multi_c_n$N.json; tail -2 $OUT/multi_c_n$N.err
timeout 600 python tools/multi_bench.py --ngpus $N --size 256 --host-size 256 --recon PARABOLIC --rk RK3 > $OUT/multi_c_c5_n$N.json 2> $OUT/multi_c_c5_n$N.err; cat $OUT/multi_c_c5_n$N.json
