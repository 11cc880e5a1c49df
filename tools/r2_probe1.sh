#!/bin/bash
# round-2 probe 1: row-major-V flash kernel parity + per-tile timelines of the transformer linears
mkdir -p gpurun_out
timeout 300 python tools/diag_attn.py > gpurun_out/r2_p1_attn.log 2>&1; echo "diag_attn rc=$?"
tail -12 gpurun_out/r2_p1_attn.log
timeout 900 python -m pytest tests/test_fgt_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/r2_p1_tests.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_p1_tests.log
for shape in "7200 1536 512 128" "7200 1960 512 128" "7200 512 1960 128" "7200 512 512 128" "7200 6272 512 128"; do
  timeout 120 python tools/trace_gemm.py $shape >> gpurun_out/r2_p1_trace.log 2>&1
done
grep -E "^linear|CTA span" gpurun_out/r2_p1_trace.log
