#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fgt_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/fold_bench.err > gpurun_out/fold_bench.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/fold_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in ('fold_unfold','fold','rownorm','gemm_tc'): print(k, d['kernels'][k])
P
