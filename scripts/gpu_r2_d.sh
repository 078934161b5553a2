#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_backward_ops_gpu.py tests/test_stock_api_gpu.py "tests/test_bench_shapes_gpu.py" -m gpu -q -rA -k "attention or gn_backward or stock or autograd or closures" > gpurun_out/tests_r2d.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/tests_r2d.log | tail -12
for cfg in "ADB_ATTN_PIPE=0 ADB_GNB_FUSED=0" "ADB_ATTN_PIPE=1 ADB_GNB_FUSED=1" "ADB_ATTN_PIPE=1 ADB_GNB_FUSED=8" "ADB_ATTN_PIPE=1 ADB_GNB_FUSED=4"; do echo "== $cfg"; env $cfg timeout 300 python scripts/prof_ops.py 256 2>&1 | grep -v "groupnorm_apply"; done | tee gpurun_out/prof_ops_r2d.log
for cfg in "ADB_ATTN_PIPE=0 ADB_GNB_FUSED=0" "ADB_ATTN_PIPE=1 ADB_GNB_FUSED=1"; do echo "== bench $cfg"; env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); kb=d['kernel_breakdown']
print('value',round(d['value'],2),'unet_only',round(d['unet_only']['value'],2),'clk',d['clocks']['sm_mhz'])
for k in ('attention','clf:attention','clf:attention_bwd','clf:groupnorm_bwd','groupnorm_apply','conv_igemm'): print(' ',k,kb[k]['ms'],kb[k].get('tflops') or kb[k].get('gbs'))
"; done | tee gpurun_out/bench_ab_r2d.log
