#!/bin/bash
# Quick GPU pass: op/backward/classifier tests + bench. usage: scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
