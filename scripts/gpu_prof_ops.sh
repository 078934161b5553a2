#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python scripts/prof_ops.py 64 > gpurun_out/prof_ops_$TAG.log 2>&1; echo "prof_ops rc=$?"; cat gpurun_out/prof_ops_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_d|attention2|gn_bwd' -s 26 -c 5 -o gpurun_out/prof_ops_$TAG -f python scripts/prof_ops.py 64 > gpurun_out/ncu_ops_$TAG.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/prof_ops_$TAG.ncu-rep
timeout 600 python -m pytest tests/test_search_gpu.py -x -q -s 2>&1 | tail -4
