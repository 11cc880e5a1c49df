#!/bin/bash
# bench.py under torchrun at N = $1 GPUs (window-DP weak scaling line + frame-sharded strong scaling legs)
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
echo "bench N=$N rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2_scale_n$N.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'n',d['n_gpus'])
print(json.dumps(d['frame_sharded'],indent=1))
P
grep -v "NCCL\|^$" gpurun_out/r2_scale_n$N.err | tail -15
