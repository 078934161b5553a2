#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_backward_ops_gpu.py -x -q -m gpu -k "gn_backward or gnb" > gpurun_out/gnb_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/gnb_tests.log
timeout 300 python scripts/prof_gnb.py 256 2>&1 | tee gpurun_out/prof_gnb_v4.log
