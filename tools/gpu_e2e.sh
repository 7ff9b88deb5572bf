#!/bin/bash
# e2e with and without NUMA binding at N GPUs: tools/gpu_e2e.sh TAG N
OUT=gpurun_out/${1:-e2e}; N=${2:-1}; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; lscpu | grep -i -E "numa|socket|model name" > $OUT/lscpu.txt
for FLAG in "" "--no-numa"; do
  if [ $N = 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"; fi
  timeout 600 $CMD --steps 6 --warmup 3 --no-cpu --no-secondary --e2e-steps 4 $FLAG > $OUT/b$N$FLAG.json 2> $OUT/b$N$FLAG.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/b$N$FLAG.json")); print("N=$N $FLAG value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("host_numa"))
except Exception as e: print("failed", e); print(open("$OUT/b$N$FLAG.err").read()[-800:])
PY
done
cat $OUT/lscpu.txt
