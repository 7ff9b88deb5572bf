#!/bin/bash
OUT=gpurun_out/${1:-curv}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gen.py tests/test_gpu_parity.py -m gpu -q -k "cylindrical or isothermal or per_step_vs_reference or ppm_on_stretched" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -25 $OUT/pytest.log | cut -c1-220
