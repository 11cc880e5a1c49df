#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, plus the
# measurements the next optimisation steps need. Run from the repository root under gpurun:
#   gpurun --timeout 600 -- 'bash tools/round2_probe.sh'
# Outputs land in gpurun_out/r2_probe_*.
set -u
mkdir -p gpurun_out
# 1. kernels not yet executed on hardware: flow splat (row a10); pipeline modes on the GPU backend
timeout 120 python -m pytest tests/test_zz_flow_warp.py -m gpu -q -rxXs 2>&1 | tail -15 > gpurun_out/r2_probe_flow_warp.log
timeout 200 python - > gpurun_out/r2_probe_modes.log 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, ".")
from fgt_b200 import pipeline as PL, synth
from tools.bench_pipeline import models
from tests.util import load_golden
for name in ("pipeline12", "pipeline_watermark", "pipeline_extrapolation"):
    g = load_golden(name); m = g["meta"]
    frames, masks = synth.pipeline_clip(seed=m["clip_seed"], N=m["N"], H=m["H"], W=m["W"])
    mode = m.get("mode", "object_removal")
    if mode == "watermark_removal":
        masks = [np.repeat(x[..., None], 3, -1) for x in masks]
    args = PL.make_args(mode=mode, imgH=m["H"], imgW=m["W"], flow_mask_dilates=m["flow_mask_dilates"], frame_dilates=m["frame_dilates"],
                        consistencyThres=m.get("consistencyThres", 5.0), H_scale=m.get("scale", 2.0), W_scale=m.get("scale", 2.0))
    hw = tuple(m.get("cfg_hw", (m["H"], m["W"])))
    be, _ = models(hw[0], hw[1], seeds=(m["fgt_seed"], m["lafc_seed"], m["raft_seed"]))
    comp = np.stack(PL.video_inpainting(frames, None if mode == "video_extrapolation" else masks, be, args))
    d = np.abs(comp.astype(np.int16) - g["comp"].astype(np.int16))
    print(name, "max level diff", int(d.max()), "mean", float(d.mean()), "frac > 2 levels", float((d > 2).mean()))
PY
# 2. where the K=512 linears lose their time: per-tile timelines of CTA 0 and CTA 147
for shape in "7200 1024 512 128" "7200 512 512 128" "7200 1960 512 128" "7200 512 1960 128"; do
  timeout 60 python tools/trace_gemm.py $shape >> gpurun_out/r2_probe_trace_gemm.log 2>&1
done
# 3. refreshed numbers of the widened rows
timeout 120 python tools/bench_poisson.py > gpurun_out/r2_probe_poisson.json 2> gpurun_out/r2_probe_poisson.err
timeout 240 python tools/bench_pipeline.py --no-cpu > gpurun_out/r2_probe_pipeline.json 2> gpurun_out/r2_probe_pipeline.err
# 4. ncu of the current LSQR kernels (two launches each)
timeout 150 ncu --set full --clock-control none --import-source on -k regex:psn_ -s 60 -c 4 -o gpurun_out/r2_probe_poisson_full -f \
  python tools/bench_poisson.py > gpurun_out/r2_probe_poisson_ncu.log 2>&1
tail -n 5 gpurun_out/r2_probe_flow_warp.log gpurun_out/r2_probe_modes.log
