#!/bin/bash
# SD-v1 GPU pass: tests, bench line, ncu launch list (batch 4), optional full-set capture of the attention / conv kernels.
# usage: scripts/gpu_sd.sh <tag> [tests|notests] [full|nofull]
TAG=${1:-x}
mkdir -p gpurun_out
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$TAG.log
fi
timeout 600 python bench.py --workload sdv1 --steps 3 --warmup 3 --dump-ops gpurun_out/ops_sd_$TAG.csv > gpurun_out/bench_sd_$TAG.json 2> gpurun_out/bench_sd_$TAG.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_sd_$TAG.json; tail -2 gpurun_out/bench_sd_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_sd_$TAG.csv \
  python bench.py --workload sdv1 --steps 1 --warmup 1 --batch 4 --no-cpu-baseline > gpurun_out/ncu_launch_sd_$TAG.log 2>&1; echo "ncu launches rc=$?"
if [ "${3:-nofull}" = "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attention_sd|conv_igemm|layernorm|geglu' -s 40 -c 30 \
    -o gpurun_out/prof_sd_$TAG -f python bench.py --workload sdv1 --steps 1 --warmup 1 --batch 8 --no-cpu-baseline > gpurun_out/ncu_full_sd_$TAG.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/prof_sd_$TAG.ncu-rep
fi
