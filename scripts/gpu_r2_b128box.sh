#!/bin/bash
# 256-wide CTA-pair tile: the CTA's 128 weight rows as ONE cta_group::2 TMA box (ADB_CONV_B128BOX=1) instead of two 64-row boxes
mkdir -p gpurun_out
ADB_CONV_B128BOX=1 timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_classifier_gpu.py -x -q -m gpu -k "conv or classifier_logits" > gpurun_out/b128_tests.log 2>&1
echo "tests with one 128-row B box rc=$?"; tail -6 gpurun_out/b128_tests.log | cut -c1-300
echo "== two 64-row boxes"; timeout 200 python scripts/prof_gnb.py 256 2>&1 | grep "256 -> 256\|512 -> 512"
echo "== one 128-row box"; ADB_CONV_B128BOX=1 timeout 200 python scripts/prof_gnb.py 256 2>&1 | grep "256 -> 256\|512 -> 512\|rror\|imeout" | head
