#!/bin/bash
# round-2 probe 3: multi-buffered accumulators, fp16 flow branch, swin_prep: parity then bench
mkdir -p gpurun_out
timeout 600 python tools/diag_gemm.py > gpurun_out/r2_p3_gemm.log 2>&1; echo "diag_gemm rc=$?"
grep -E "FAIL|FAILURES|Error|error" gpurun_out/r2_p3_gemm.log | head
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_fgt_gpu.py tests/test_zz_flow_warp.py -q -m gpu > gpurun_out/r2_p3_tests.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_p3_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_p3_bench.json 2> gpurun_out/r2_p3_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_p3_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('roofline',{k:d['roofline'][k] for k in ('kernel','achieved','frac','modules','worst_module')})
for k,v in d['kernels'].items(): print(k,v)
for k,v in d['modules'].items(): print(k,v)
P
tail -5 gpurun_out/r2_p3_bench.err
timeout 300 python tools/profile_layers.py > gpurun_out/r2_p3_layers.log 2>&1; tail -80 gpurun_out/r2_p3_layers.log
