#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/microbench/tmem_mufu_mix | tee gpurun_out/tmem_mufu_mix.log
python scripts/prof_one.py 64 attn_bwd > gpurun_out/plain_r2h.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_fused' -o gpurun_out/prof_r2h -f python scripts/prof_one.py 64 attn_bwd > gpurun_out/ncu_r2h.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2h.log
