#!/bin/bash
# Round-2 first pass: GPU suite with captured prints (measured values for threshold tightening), the smoke-outlier
# diagnosis, a baseline bench line, and compute-sanitizer passes over the small op tests.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA > gpurun_out/r2_tests_rA.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_rA.log
timeout 300 python scripts/diag_outlier.py > gpurun_out/r2_diag_outlier.log 2>&1; echo "diag rc=$?"; cat gpurun_out/r2_diag_outlier.log
timeout 600 python bench.py --steps 5 --warmup 3 --dump-ops gpurun_out/ops_r2_base.csv > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r2_base.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_ops_gpu.py tests/test_backward_ops_gpu.py -m gpu -x -q -k "conv3x3 or gemm_k1 or attention or gn_backward" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_ops_gpu.py tests/test_backward_ops_gpu.py -m gpu -x -q -k "gemm_k1 or (attention and 64) or (gn_backward and 8)" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_racecheck.log
