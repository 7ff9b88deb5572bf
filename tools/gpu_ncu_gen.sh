#!/bin/bash
# ncu --set full of the general-path kernels of the LDW bench: tools/gpu_ncu_gen.sh TAG
OUT=gpurun_out/${1:-ncugen}; mkdir -p $OUT
PB200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gen_vgrad|gen_sweep|gen_finish" --launch-skip 10 -c 5 \
  -o $OUT/full_ldw -f python bench.py --workload ldw --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > $OUT/full.log 2>&1
ncu -i $OUT/full_ldw.ncu-rep --page raw --csv > $OUT/full_ldw_raw.csv 2>/dev/null
ls -la $OUT
