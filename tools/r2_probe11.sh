#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fgt_gpu.py tests/test_clip.py tests/test_frame_shard_gpu.py -q -m gpu -x > gpurun_out/r2_p11_tests.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_p11_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
FGT_PDL=0 timeout 300 python tools/profile_small.py > gpurun_out/r2_p11_small.log 2>&1; head -1 gpurun_out/r2_p11_small.log
