#!/bin/bash
# 11 candidates on 4 ranks = 8 whole + 3 batch-sharded tail candidates (three deferred all-reduces); FIDs must equal the
# single-GPU run's. usage: scripts/gpu_r2_n4_tail.sh
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 \
  scripts/population_eval.py --candidates 11 --num_samples 1000 --batch_size 256 --guided > gpurun_out/pop_n4_tail3.json 2> gpurun_out/pop_n4_tail3.err
echo "N=4 rc=$?"; tail -2 gpurun_out/pop_n4_tail3.err
timeout 600 python scripts/population_eval.py --candidates 11 --num_samples 1000 --batch_size 256 --guided > gpurun_out/pop_n1_11cand.json 2> gpurun_out/pop_n1_11cand.err
echo "N=1 rc=$?"; tail -2 gpurun_out/pop_n1_11cand.err
python - <<'PY'
import json
a=json.loads(open("gpurun_out/pop_n4_tail3.json").read().strip().splitlines()[-1])
b=json.loads(open("gpurun_out/pop_n1_11cand.json").read().strip().splitlines()[-1])
fa,fb=a["fids"],b["fids"]
print("N=4", {k:v for k,v in a.items() if k!="fids"})
print("N=1", {k:v for k,v in b.items() if k!="fids"})
print("max |dFID|", max(abs(x-y) for x,y in zip(fa,fb)), "max rel", max(abs(x-y)/abs(y) for x,y in zip(fa,fb)))
PY
