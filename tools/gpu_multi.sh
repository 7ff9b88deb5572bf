#!/bin/bash
# Multi-GPU visit: slab/NCCL parity test, weak scaling of C2 and C5 at N = 1 and $1.
N=${1:-2}
OUT=gpurun_out/multi$N
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "slab_nccl" > $OUT/pytest.log 2>&1; tail -n 2 $OUT/pytest.log
for cfg in "c2" "c5 --size 256 --recon PARABOLIC --rk RK3"; do
  set -- $cfg; name=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu "$@" > $OUT/${name}_1.json 2> $OUT/${name}_1.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-cpu "$@" > $OUT/${name}_$N.json 2> $OUT/${name}_$N.err
  python - <<PY
import json
a=json.load(open("$OUT/${name}_1.json")); b=json.load(open("$OUT/${name}_$N.json"))
print("$name", "N=1 %.3f ms  N=$N %.3f ms  efficiency %.3f" % (a["ms_per_step"], b["ms_per_step"], b["value"]/($N*a["value"])))
PY
done
