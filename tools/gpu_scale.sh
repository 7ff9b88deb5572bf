#!/bin/bash
# Multi-GPU visit: weak scaling of C2 at N = $1 (and the 2-GPU slab parity test), C3 RT at N = $1
N=${1:-8}
OUT=gpurun_out/scale$N
mkdir -p $OUT
run() {  # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json")); print("$name", "N=%d"%d["n_gpus"], round(d["ms_per_step"],3), "ms/step", round(d["value"],1), d["unit"], d["config"]["workload"])
except Exception as e: print("$name failed", e, open("$OUT/$name.err").read()[-800:])
PY
}
run c2 --steps 20 --warmup 3 --no-e2e --no-cpu
run c3_rt --workload rt --size ${2:-1024} --steps 20 --warmup 3 --no-e2e --no-cpu
run c5 --size 256 --recon PARABOLIC --rk RK3 --steps 20 --warmup 3 --no-e2e --no-cpu

