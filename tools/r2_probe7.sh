#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_morph.py tests/test_kernels_gpu.py tests/test_fgt_gpu.py tests/test_pipeline.py -q -m gpu -x > gpurun_out/r2_p7_tests.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_p7_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_p7_bench.json 2> gpurun_out/r2_p7_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_p7_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k,v in d['kernels'].items(): print(k,v)
for k,v in d['modules'].items(): print(k,v)
P
timeout 600 python bench.py --config 4 --steps 8 --warmup 3 > gpurun_out/r2_p7_c4.json 2> gpurun_out/r2_p7_c4.err; echo "c4 rc=$?"; tail -c 1200 gpurun_out/r2_p7_c4.json; tail -3 gpurun_out/r2_p7_c4.err
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2_p7_c3.json 2> gpurun_out/r2_p7_c3.err; echo "c3 rc=$?"; tail -c 1500 gpurun_out/r2_p7_c3.json; tail -3 gpurun_out/r2_p7_c3.err
timeout 600 python tools/bench_pipeline.py > gpurun_out/r2_p7_pipeline.json 2> gpurun_out/r2_p7_pipeline.err; echo "pipeline rc=$?"; tail -c 1500 gpurun_out/r2_p7_pipeline.json
