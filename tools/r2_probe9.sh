#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_kernels_gpu.py tests/test_fgt_gpu.py tests/test_clip.py tests/test_frame_shard_gpu.py tests/test_pipeline.py -q -m gpu --durations=5 > gpurun_out/r2_gpu_tests_b.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_gpu_tests_b.log
timeout 1200 python bench.py > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_line.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'sched',d['driver_schedule']['value'])
print('roofline',{k:d['roofline'][k] for k in ('kernel','achieved','frac','modules','worst_module','traffic')})
print('eager',d['gpu_eager_baseline']); print('cpu',d['cpu_baseline']); print('clocks', d['clocks'])
for k,v in d['kernels'].items(): print(k,v)
for k,v in d['modules'].items(): print(k,v)
P
tail -3 gpurun_out/r2_bench.err
timeout 300 python tools/profile_layers.py > gpurun_out/r2_layers.log 2>&1; tail -3 gpurun_out/r2_layers.log
timeout 900 python bench.py --config 5 --steps 8 --warmup 3 > gpurun_out/r2_c5.json 2> gpurun_out/r2_c5.err; echo "c5 rc=$?"; tail -c 1500 gpurun_out/r2_c5.json; tail -3 gpurun_out/r2_c5.err
