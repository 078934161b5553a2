#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_gemm_gnbwd.py 256 time 2>&1 | tee gpurun_out/r02_gemm_gnbwd_times.txt
timeout 120 python scripts/prof_gemm_gnbwd.py 256 once > /dev/null && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|gn_bwd" -o gpurun_out/r02_gemm_gnbwd -f python scripts/prof_gemm_gnbwd.py 256 once > gpurun_out/ncu_small.log 2>&1
echo "ncu rc=$?"
