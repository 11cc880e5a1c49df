#!/bin/bash
# last sanity pass of the shipped tree: smoke(), the reference arm, a short default line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/last_ref.err | tee gpurun_out/last_ref.json | cut -c1-600
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/last_bench.err | tee gpurun_out/last_bench.json | cut -c1-400
