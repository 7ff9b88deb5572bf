#!/bin/bash
OUT=gpurun_out/${1:-cool}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_glibc_math.py tests/test_gpu_gen.py tests/test_dropin_gpu.py -m gpu -x -q -k "libm or cool or blondin or line_driven or ldw" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest.log
