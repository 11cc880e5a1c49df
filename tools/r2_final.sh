#!/bin/bash
# final evidence of the round on one GPU: full GPU suite, full default bench line, per-layer profile, config 4 at N=1,
# and a fresh ncu metric capture of one forward
mkdir -p gpurun_out gpurun_out/ncu_r2
timeout 2400 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r2_gpu_tests.log
timeout 1200 python bench.py > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_line.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'serial',d['e2e']['serial_value'],'launches',d['gpu_launches'],'sched',d['driver_schedule'])
print('roofline',{k:d['roofline'][k] for k in ('kernel','achieved','frac','modules','worst_module')})
print('eager',d['gpu_eager_baseline']); print('cpu',d['cpu_baseline']); print('clocks', d['clocks'])
for k,v in d['kernels'].items(): print(k,v)
for k,v in d['modules'].items(): print(k,v)
P
tail -3 gpurun_out/r2_bench.err
timeout 300 python tools/profile_layers.py > gpurun_out/r2_layers.log 2>&1; tail -2 gpurun_out/r2_layers.log
timeout 600 python bench.py --config 4 --steps 8 --warmup 3 > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err; echo "c4 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2_c4_n1.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('metric','value','n_gpus','ms_per_step')}, d['e2e']['value'])"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
L=$(python -c "import json;d=json.loads(open('gpurun_out/r2_bench_line.json').read().strip().splitlines()[-1]);print(d['gpu_launches']//d['steps'])")
echo "launches per forward: $L" | tee gpurun_out/ncu_r2/launches_per_forward.txt
FGT_BENCH_GRAPH=0 timeout 900 ncu --clock-control none --kernel-name-base demangled --csv --metrics $M -k regex:fgt:: -s $((3*L)) -c $L --log-file gpurun_out/ncu_r2/model_metrics.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_r2/model_bench.out 2> gpurun_out/ncu_r2/model_bench.err
echo "ncu model rc=$?"
