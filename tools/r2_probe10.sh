#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_fgt_gpu.py tests/test_clip.py tests/test_frame_shard_gpu.py -q -m gpu -x > gpurun_out/r2_p10_tests.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_p10_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_p10_bench.json 2> gpurun_out/r2_p10_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_p10_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'sched',d['driver_schedule']['value'])
print({k:v for k,v in d['roofline']['modules'].items()})
for k,v in d['kernels'].items(): print(k,v)
P
tail -3 gpurun_out/r2_p10_bench.err
timeout 300 python tools/profile_layers.py > gpurun_out/r2_p10_layers.log 2>&1; grep -E "enc14|\.o |ffn2|sum of" gpurun_out/r2_p10_layers.log | head -20
