#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_population_gpu.py -m gpu -q -rA > gpurun_out/tests_r2k.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|population of" gpurun_out/tests_r2k.log | tail -6; grep -E "^E  " gpurun_out/tests_r2k.log | head -8
timeout 900 python scripts/population_eval.py --candidates 6 --num_samples 1000 --batch_size 256 --guided > gpurun_out/pop_n1_r2k.json 2> gpurun_out/pop_n1_r2k.err; echo "pop rc=$?"; cut -c1-900 gpurun_out/pop_n1_r2k.json; tail -3 gpurun_out/pop_n1_r2k.err
