#!/bin/bash
# drop-in + multi-GPU C layer tests
OUT=gpurun_out/${1:-dropin}; mkdir -p $OUT
timeout 1700 python -m pytest tests/test_dropin_gpu.py tests/test_multi_gpu.py -m gpu -x -q --durations=8 ${2:+-k "$2"} > $OUT/pytest.log 2>&1
echo "exit $?"; tail -25 $OUT/pytest.log
