#!/bin/bash
# packed f32x2 softmax arithmetic in the attention kernels: tests, kernel timings and bench A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_backward_ops_gpu.py tests/test_sd_unet_gpu.py tests/test_classifier_gpu.py -x -q -m gpu -k "attention or attn or sd or guided or config1" > gpurun_out/f32x2_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/f32x2_tests.log
P=$PWD/autodiffusion_b200/lib/libadb200_prev.so
echo "== prev"; ADB_LIB_PATH=$P timeout 300 python scripts/prof_ops.py 256 2>&1 | grep attention
echo "== new"; timeout 300 python scripts/prof_ops.py 256 2>&1 | grep attention
for rep in 1 2; do
  ADB_LIB_PATH=$P timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-roofline > gpurun_out/f32x2_prev_$rep.json 2> gpurun_out/f32x2_prev_$rep.err
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-roofline > gpurun_out/f32x2_new_$rep.json 2> gpurun_out/f32x2_new_$rep.err
done
python - <<'PY'
import json
for n in ("prev_1","new_1","prev_2","new_2"):
    try:
        d=json.loads(open(f"gpurun_out/f32x2_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],2), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "ERR", e)
PY
