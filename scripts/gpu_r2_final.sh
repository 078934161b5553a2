#!/bin/bash
# Round-2 final measurements: default bench line (all extras), ncu launch list of the same command (batch 32), ncu --set full
# of the conv set at batch 256 (roofline.traffic source), per-op dump.
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 --dump-ops gpurun_out/ops_r2_final.csv > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_r2_final.json; tail -2 gpurun_out/bench_r2_final.err
python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-roofline --no-extras > gpurun_out/plain_launches.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-roofline --no-extras > gpurun_out/ncu_launch_r2.log 2>&1; echo "ncu launches rc=$?"
python scripts/prof_one.py 256 convset > gpurun_out/plain_convset.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -o gpurun_out/prof_r2_convset -f python scripts/prof_one.py 256 convset > gpurun_out/ncu_convset.log 2>&1; echo "ncu convset rc=$?"; tail -2 gpurun_out/ncu_convset.log
