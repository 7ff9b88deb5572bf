#!/bin/bash
OUT=gpurun_out/${1:-iso}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gen.py tests/test_dropin_gpu.py -m gpu -x -q -k "isothermal" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest.log | cut -c1-250
