#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_fgt_gpu.py -q -m gpu -x > gpurun_out/r2_p5_tests.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_p5_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_p5_bench.json 2> gpurun_out/r2_p5_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_p5_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'sched', d['driver_schedule']['value'] if d['driver_schedule'] else None)
print('roofline',{k:d['roofline'][k] for k in ('kernel','achieved','frac','modules','worst_module')})
for k,v in d['kernels'].items(): print(k,v)
for k,v in d['modules'].items(): print(k,v)
P
tail -5 gpurun_out/r2_p5_bench.err
timeout 300 python tools/profile_layers.py > gpurun_out/r2_p5_layers.log 2>&1; grep -E "dec|pos|gate|dwconv|conv_tail|sum of" gpurun_out/r2_p5_layers.log | tail -30
