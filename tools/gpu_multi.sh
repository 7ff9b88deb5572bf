#!/bin/bash
# Multi-GPU visit: slab/NCCL parity test, drop-in tests, weak-scaling bench at N = 1 and $1.
N=${1:-2}
OUT=gpurun_out/multi$N
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "slab_nccl or dropin" -rs > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $OUT/bench1.json 2> $OUT/bench1.err
cat $OUT/bench1.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-cpu > $OUT/bench$N.json 2> $OUT/bench$N.err
cat $OUT/bench$N.json | cut -c1-300; tail -3 $OUT/bench$N.err
