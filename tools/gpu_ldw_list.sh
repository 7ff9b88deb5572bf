#!/bin/bash
# ncu launch list of the LDW bench (fused form; FUSED="1 0" adds the unfused one): tools/gpu_ldw_list.sh TAG
OUT=gpurun_out/${1:-ldwlist}; mkdir -p $OUT
for F in ${FUSED:-1}; do
  PB200_GEN_FUSED=$F PB200_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_f$F.csv python bench.py --workload ldw --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > $OUT/l$F.log 2>&1
  tail -1 $OUT/l$F.log | cut -c1-200
done
