#!/bin/bash
# The driver's N-GPU command (bench.py under torchrun): at N = 8 the `population_eval` key is BASELINE configs[2] itself,
# 50 candidates x 1000 samples. usage: scripts/gpu_r2_n8.sh <N>
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err
echo "bench N=$N rc=$?"; cut -c1-260 gpurun_out/bench_n${N}_r2.json; tail -3 gpurun_out/bench_n${N}_r2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_r2.json").read().strip().splitlines()[-1])
    pe=dict(d.get("population_eval") or {}); pe.pop("fids",None)
    print("value",d["value"],"e2e",d["e2e"]["value"],"population_eval",json.dumps(pe))
except Exception as e:
    print("ERR",e)
PY
