#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 scripts/population_eval.py --candidates 5 --num_samples 1000 --batch_size 256 --guided > gpurun_out/pop_n2_r2.json 2> gpurun_out/pop_n2_r2.err; echo "pop n2 rc=$?"; cut -c1-1000 gpurun_out/pop_n2_r2.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/pop_n2_r2.err | tail -5
timeout 600 python scripts/population_eval.py --candidates 5 --num_samples 1000 --batch_size 256 --guided > gpurun_out/pop_n1_5c_r2.json 2>/dev/null; cut -c1-400 gpurun_out/pop_n1_5c_r2.json; python -c "
import json
a=json.loads(open('gpurun_out/pop_n2_r2.json').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/pop_n1_5c_r2.json').read().strip().splitlines()[-1])
print('fid checksum n2',a['fid_checksum'],'n1',b['fid_checksum'],'first3',a['fid_first3'],b['fid_first3'])"
