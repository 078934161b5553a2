#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_population_gpu.py -m gpu -q 2>&1 | tail -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_r2.json 2> gpurun_out/bench_n2_r2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_r2.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'four_step',d['four_step']['value'],'stock',d['stock_api']['value'])
print(json.dumps(d['population_eval'])[:900])
PY
tail -3 gpurun_out/bench_n2_r2.err
