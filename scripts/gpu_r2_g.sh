#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_ops_gpu.py tests/test_bench_shapes_gpu.py tests/test_classifier_gpu.py -m gpu -q -rA -x -k "attention_backward or classifier_logits or guided_sampling" > gpurun_out/tests_r2g.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|attention_backward" gpurun_out/tests_r2g.log | tail -22
grep -E "^E  |timeout|Error" gpurun_out/tests_r2g.log | head -10
for f in 0 1; do echo "== ADB_ATTN_BWD_FUSED=$f"; ADB_ATTN_BWD_FUSED=$f timeout 300 python scripts/prof_ops.py 256 2>&1 | grep "attention"; done | tee gpurun_out/prof_attn_r2g.log
