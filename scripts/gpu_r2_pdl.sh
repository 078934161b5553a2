#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_ops_gpu.py tests/test_backward_ops_gpu.py tests/test_classifier_gpu.py tests/test_unet_gpu.py -m gpu -q -x > gpurun_out/tests_pdl.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_pdl.log
for f in 1 0; do echo "== bench ADB_NO_PDL=$f"; ADB_NO_PDL=$f timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'unet_only',round(d['unet_only']['value'],2),'clk',d['clocks']['sm_mhz'])
"; done | tee gpurun_out/bench_ab_pdl.log
for f in 1 0; do echo "== bench again ADB_NO_PDL=$f"; ADB_NO_PDL=$f timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'unet_only',round(d['unet_only']['value'],2),'clk',d['clocks']['sm_mhz'])
"; done | tee -a gpurun_out/bench_ab_pdl.log
