#!/bin/bash
# GPU suite with captured prints. usage: scripts/gpu_r2_tests.sh <tag> [pytest args...]
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rA "$@" > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/tests_$TAG.log | tail -15
