#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_stock_api_gpu.py tests/test_bench_shapes_gpu.py tests/test_backward_ops_gpu.py -m gpu -q -rA > gpurun_out/tests_r2c.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/tests_r2c.log | tail -12
for pipe in 0 1; do echo "== ADB_ATTN_BWD_PIPE=$pipe"; ADB_ATTN_BWD_PIPE=$pipe timeout 300 python scripts/prof_ops.py 256 2>&1 | tee gpurun_out/prof_ops_pipe$pipe.log; done
timeout 900 python bench.py --steps 5 --warmup 3 --dump-ops gpurun_out/ops_r2c.csv > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
