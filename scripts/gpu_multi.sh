#!/bin/bash
# Multi-GPU pass (N ranks on one box): bench line at N, population eval at N. usage: scripts/gpu_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
echo "bench N=$N rc=$?"; cut -c1-300 gpurun_out/bench_n${N}_$TAG.json; tail -3 gpurun_out/bench_n${N}_$TAG.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  scripts/population_eval.py --candidates 8 --num_samples 1000 --batch_size 125 --feature_dim 2048 --guided > gpurun_out/pop_n${N}_$TAG.json 2> gpurun_out/pop_n${N}_$TAG.err
echo "pop N=$N rc=$?"; cat gpurun_out/pop_n${N}_$TAG.json; tail -3 gpurun_out/pop_n${N}_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --workload sdv1 --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sd_n${N}_$TAG.json 2> gpurun_out/bench_sd_n${N}_$TAG.err
echo "sd bench N=$N rc=$?"; cut -c1-300 gpurun_out/bench_sd_n${N}_$TAG.json; tail -3 gpurun_out/bench_sd_n${N}_$TAG.err
