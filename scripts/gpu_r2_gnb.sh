#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_backward_ops_gpu.py -x -q -m gpu -k "gn_backward or gnb" > gpurun_out/gnb_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/gnb_tests.log
timeout 300 python scripts/prof_gnb.py 256 2>&1 | tee gpurun_out/prof_gnb_v3.log
for rep in 1 2; do
  ADB_NO_GNB_FUSE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/gnb_off_$rep.json 2> gpurun_out/gnb_off_$rep.err
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/gnb_on_$rep.json 2> gpurun_out/gnb_on_$rep.err
done
python - <<'PY'
import json
for n in ("off_1","on_1","off_2","on_2"):
    try:
        d=json.loads(open(f"gpurun_out/gnb_{n}.json").read().strip().splitlines()[-1])
        kb={k:v["ms"] for k,v in d.get("kernel_breakdown",{}).items() if k in ("conv_igemm","clf:groupnorm_bwd")}
        print(n, round(d["value"],2), round(d["ms_per_step"],1), kb)
    except Exception as e:
        print(n, "ERR", e)
PY


