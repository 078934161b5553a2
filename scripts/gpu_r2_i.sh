#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rA -x > gpurun_out/tests_r2i.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/tests_r2i.log | tail -8
grep -E "^E  " gpurun_out/tests_r2i.log | head -8
for f in 0 1; do echo "== bench ADB_ATTN_BWD_FUSED=$f"; ADB_ATTN_BWD_FUSED=$f timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); kb=d['kernel_breakdown']
print('value',round(d['value'],2),'unet_only',round(d['unet_only']['value'],2),'clk',d['clocks']['sm_mhz'])
for k in ('clf:attention_bwd','clf:attention','attention','clf:groupnorm_bwd','conv_igemm'): print(' ',k,kb[k]['ms'],kb[k].get('tflops') or kb[k].get('gbs'))
"; done | tee gpurun_out/bench_ab_r2i.log
