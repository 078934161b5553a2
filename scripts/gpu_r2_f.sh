#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_inception_gpu.py tests/test_stock_api_gpu.py tests/test_backward_ops_gpu.py -m gpu -q -rA > gpurun_out/tests_r2f.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|pool_3|resize 64" gpurun_out/tests_r2f.log | tail -20
grep -E "^E  " gpurun_out/tests_r2f.log | head -20
