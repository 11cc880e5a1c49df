#!/bin/bash
# round-2 probe 2: TMA-store epilogue + 1-term flow branch: parity, then timings
mkdir -p gpurun_out
timeout 600 python tools/diag_gemm.py > gpurun_out/r2_p2_gemm.log 2>&1; echo "diag_gemm rc=$?"
grep -E "FAIL|FAILURES|Error|error" gpurun_out/r2_p2_gemm.log | head
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_fgt_gpu.py -x -q -m gpu > gpurun_out/r2_p2_tests.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_p2_tests.log
rm -f gpurun_out/r2_p2_trace.log
for shape in "7200 1536 512 128" "7200 1960 512 128" "7200 6272 512 128"; do
  timeout 120 python tools/trace_gemm.py $shape >> gpurun_out/r2_p2_trace.log 2>&1
done
grep -E "^linear|CTA span" gpurun_out/r2_p2_trace.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_p2_bench.json 2> gpurun_out/r2_p2_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_p2_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('roofline',{k:d['roofline'][k] for k in ('kernel','achieved','frac','modules','worst_module')})
print('eager',d['gpu_eager_baseline'])
print('cpu',d['cpu_baseline'])
for k,v in d['kernels'].items(): print(k,v)
for k,v in d['modules'].items(): print(k,v)
P
tail -5 gpurun_out/r2_p2_bench.err
