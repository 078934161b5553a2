#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench line, ncu launch list (batch 32). Outputs under gpurun_out/.
# usage: scripts/gpu_round.sh <tag> [tests|notests]
TAG=${1:-x}
mkdir -p gpurun_out
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$TAG.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 --dump-ops gpurun_out/ops_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"; tail -c 300 gpurun_out/ncu_launch_$TAG.log
# one `ncu --set full` capture of the classifier-guidance kernels + the conv kernel (batch 32: 40 replays per launch)
if [ "${3:-nofull}" = "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd|gn_bwd|conv_igemm|attention2' -s 700 -c 24 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
fi
