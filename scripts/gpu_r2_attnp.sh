#!/bin/bash
# persistent attention forward (ADB_ATTN_PERSIST=1): parity tests, kernel timings, bench A/B
mkdir -p gpurun_out
ADB_ATTN_PERSIST=1 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_backward_ops_gpu.py tests/test_bench_shapes_gpu.py -x -q -m gpu -k "attention" > gpurun_out/attnp_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/attnp_tests.log | cut -c1-300
echo "== per-CTA kernel"; timeout 100 python scripts/prof_ops.py 256 2>&1 | grep attention | cut -c1-64
echo "== persistent"; ADB_ATTN_PERSIST=1 timeout 100 python scripts/prof_ops.py 256 2>&1 | grep "attention\|rror\|imeout" | cut -c1-64
