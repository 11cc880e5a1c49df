#!/bin/bash
# bench.py --config 4 (T=80 clip, 16 windows sharded over the ranks) under torchrun at N = $1 GPUs
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --config 4 --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_c4_n$N.json 2> gpurun_out/r2_c4_n$N.err
echo "c4 N=$N rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2_c4_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','n_gpus','ms_per_step','scaling')}, d['e2e']['value'], d['detail'])
P
grep -v "NCCL\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_c4_n$N.err | tail -8
