#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_morph.py tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/r2_p6_tests.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_p6_tests.log
FGT_PDL=0 timeout 300 python tools/profile_small.py --layers > gpurun_out/r2_p6_small_pdl0.log 2>&1
FGT_PDL=1 timeout 300 python tools/profile_small.py > gpurun_out/r2_p6_small_pdl1.log 2>&1
head -1 gpurun_out/r2_p6_small_pdl0.log; head -1 gpurun_out/r2_p6_small_pdl1.log
sort -k3 -n -r gpurun_out/r2_p6_small_pdl0.log | head -30; tail -1 gpurun_out/r2_p6_small_pdl0.log
