#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --config 4 --steps 8 --warmup 3 > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err; echo "c4 rc=$?"
timeout 500 python bench.py --config 5 --steps 8 --warmup 3 > gpurun_out/r2_c5.json 2> gpurun_out/r2_c5.err; echo "c5 rc=$?"
tail -3 gpurun_out/r2_c4_n1.err gpurun_out/r2_c5.err
python - <<'P'
import json
for f in ('r2_c4_n1','r2_c5'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e'])
P
