#!/bin/bash
# validation of the division-free im2col staging / dwconv and a bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fgt_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/div_tests.log
timeout 300 python tools/profile_layers.py 2>&1 | grep -E "im2col|dwconv|rownorm|TOTAL|total" | head -30 > gpurun_out/div_layers.log
timeout 400 python bench.py --no-cpu-baseline --no-eager-baseline > gpurun_out/div_bench.json 2> gpurun_out/div_bench.err
cat gpurun_out/div_tests.log gpurun_out/div_layers.log; python -c "
import json; d=json.loads(open('gpurun_out/div_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
