#!/bin/bash
# Round-2 closing run: whole GPU suite, smoke, the default bench line (all extras) with the per-op dump, and the ncu launch
# list of the same command at batch 32.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/tests_r2_final2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r2_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2_final2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2_final2.log
timeout 1500 python bench.py --steps 5 --warmup 3 --dump-ops gpurun_out/ops_r2_final2.csv > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_r2_final2.json; tail -2 gpurun_out/bench_r2_final2.err
python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-roofline --no-extras > gpurun_out/plain_launches.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_final2.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-roofline --no-extras > gpurun_out/ncu_launch_r2.log 2>&1; echo "ncu launches rc=$?"
