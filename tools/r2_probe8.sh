#!/bin/bash
# full GPU suite + full default bench line + per-layer profile (final numbers of the round)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -16 gpurun_out/r2_gpu_tests.log
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
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2_c3.json 2> gpurun_out/r2_c3.err; echo "c3 rc=$?"; tail -c 1800 gpurun_out/r2_c3.json; tail -3 gpurun_out/r2_c3.err
timeout 900 python bench.py --config 5 --steps 8 --warmup 3 > gpurun_out/r2_c5.json 2> gpurun_out/r2_c5.err; echo "c5 rc=$?"; tail -c 1500 gpurun_out/r2_c5.json; tail -3 gpurun_out/r2_c5.err
