#!/bin/bash
# UNet forward and classifier guidance of a DDIM step on two streams (ADB_CONCURRENT_GUIDANCE=1): parity tests, then bench A/B
mkdir -p gpurun_out
ADB_CONCURRENT_GUIDANCE=1 timeout 900 python -m pytest tests/test_classifier_gpu.py tests/test_bench_shapes_gpu.py tests/test_stock_api_gpu.py tests/test_population_gpu.py -x -q -m gpu > gpurun_out/conc_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/conc_tests.log
for rep in 1 2; do
  for c in 0 1; do
    ADB_CONCURRENT_GUIDANCE=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-roofline > gpurun_out/conc_${c}_$rep.json 2> gpurun_out/conc_${c}_$rep.err
  done
done
python - <<'PY'
import json
for n in ("0_1","1_1","0_2","1_2"):
    try:
        d=json.loads(open(f"gpurun_out/conc_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],2), round(d["ms_per_step"],1), round(d["e2e"]["value"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "ERR", e)
PY
