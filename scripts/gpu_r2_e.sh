#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 2 4; do echo "== ADB_GNB_FUSED=$f"; ADB_GNB_FUSED=$f timeout 300 python scripts/prof_ops.py 256 2>&1 | grep gn_backward; done | tee gpurun_out/prof_gnb_r2e.log
ADB_GNB_FUSED=0 python scripts/prof_one.py 64 attn,attn_bwd,gn_bwd,gn,gemm > gpurun_out/plain_r2e.log 2>&1 && \
ADB_GNB_FUSED=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'attention2|attn_bwd|gn_bwd_kernel|gn_apply|conv_igemm' -o gpurun_out/prof_r2e -f python scripts/prof_one.py 64 attn,attn_bwd,gn_bwd,gn,gemm > gpurun_out/ncu_r2e.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2e.log; ls -la gpurun_out/prof_r2e.ncu-rep
