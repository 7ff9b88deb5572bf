#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the sweep kernels.
# usage: tools/gpu_check.sh TAG [skip-tests]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest.log
  tail -5 $OUT/pytest.log
fi
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; cat $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_ --launch-skip 12 -c 4 \
  -o $OUT/full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/full.log 2>&1
ncu -i $OUT/full.ncu-rep --page raw --csv > $OUT/full_raw.csv 2>/dev/null
ls -la $OUT
# general path (LDW workload): full capture of one stage worth of gen_* kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gen_ --launch-skip 96 -c 16 \
  -o $OUT/full_ldw -f python bench.py --workload ldw --steps 2 --warmup 3 > $OUT/full_ldw.log 2>&1
ncu -i $OUT/full_ldw.ncu-rep --page raw --csv > $OUT/full_ldw_raw.csv 2>/dev/null
rm -f $OUT/full_ldw.ncu-rep $OUT/full.ncu-rep
ls -la $OUT
# the line-driven-wind drop-in once more with the library's table readers behind read_sirocco_fluxes
PB200_FAST_TABLES=1 timeout 600 python -m pytest tests/test_dropin_gpu.py -q -k line_driven > $OUT/pytest_fast_tables.log 2>&1
tail -2 $OUT/pytest_fast_tables.log
