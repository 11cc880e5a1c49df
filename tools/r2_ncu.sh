#!/bin/bash
# round-2 ncu evidence (1 GPU; never under a multi-rank command). Outputs under gpurun_out/ncu_r2/, summarised by
# tools/summarize_ncu2.py into profiles/.
set -u
O=gpurun_out/ncu_r2
mkdir -p $O
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
NCU="ncu --clock-control none --kernel-name-base demangled --csv"
# launches per eager forward, from a plain run
L=$(FGT_BENCH_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['gpu_launches']//2)")
echo "launches per forward: $L" | tee $O/launches_per_forward.txt
# (1) one bench.py step: every launch with hardware metrics (eager launches so that one step = L launches)
FGT_BENCH_GRAPH=0 timeout 1500 $NCU --metrics $M -k regex:fgt:: -s $((3*L)) -c $L --log-file $O/model_metrics.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/model_bench.out 2> $O/model_bench.err
echo "model rc=$?"
# (2) the other rows' kernels
for t in raft lafc prop fill poisson splat tail; do
  timeout 900 $NCU --metrics $M -k regex:fgt:: -c 600 --log-file $O/${t}_metrics.csv python tools/ncu_targets.py $t > $O/${t}.out 2>&1
  echo "$t rc=$?"
done
# (3) --set full of the top kernels (source-level): gemm_tc on the transformer linears + an encoder conv, flash, conv_tail, swin_prep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 9 -o $O/gemm_full -f python tools/ncu_targets.py gemm > $O/gemm_full.out 2>&1; echo "gemm_full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_kernel -c 2 -o $O/flash_full -f python tools/ncu_targets.py flash > $O/flash_full.out 2>&1; echo "flash_full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tail -c 1 -s 1 -o $O/tail_full -f python tools/ncu_targets.py tail > $O/tail_full.out 2>&1; echo "tail_full rc=$?"
ls -la $O
