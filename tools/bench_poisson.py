"""Stage timing for Poisson blending at BASELINE config 2's clip shape: the 10 frames of a
240x432 clip blended as one batch (30 LSQR systems) — host-to-host wall time, device time of the iteration
kernels (CUDA events), iterations, achieved bytes/s of the two streaming kernels, and the CPU oracle timed on a
bounded sample (one frame, one channel) next to it. Prints one JSON object. Run under gpurun."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, poisson as P, synth  # noqa: E402
from oracle import poisson_oracle as PO  # noqa: E402

F, H, W = 10, 240, 432
trg, gx, gy, hole, gm = synth.poisson_inputs(seed=21, F=F, H=H, W=W)
hole[-1], gm[-1] = hole[0], gm[0]            # the synthetic set leaves its last hole empty; reuse frame 0's
trg[-1][hole[-1]] = 0
res = dict(frames=F, H=H, W=W, hole_fraction=float(hole.mean()))

for _ in range(2):
    out, unf, istop, itn = P.poisson_blend_batch(trg, gx, gy, hole, gm, return_info=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    o = P.poisson_blend_batch(trg, gx, gy, hole, gm)
    o[0].cpu(), o[1].cpu()
res["ms_per_clip_host_to_host"] = (time.perf_counter() - t0) / 3 * 1e3
res["lsqr_iterations"] = dict(min=int(itn.min()), max=int(itn.max()), istop=sorted(set(istop.flatten().tolist())))

# device time of the iteration kernels alone: inputs resident, CUDA events around CHUNK-sized calls
dev = torch.device("cuda:0")
t64 = lambda a: torch.as_tensor(a).to(dev, torch.float64).contiguous()
a_trg, a_gx, a_gy = t64(trg), t64(gx), t64(gy)
a_hole, a_gm = torch.as_tensor(hole).to(dev), torch.as_tensor(gm).to(dev)
torch.cuda.synchronize()
before = lib.COUNTERS["launches"]
P.poisson_blend_batch(a_trg, a_gx, a_gy, a_hole, a_gm)
launches = lib.COUNTERS["launches"] - before
for mode, key in ((True, "ms_per_clip_device_resident"), (False, "ms_per_clip_device_resident_stream_launches")):
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        P.poisson_blend_batch(a_trg, a_gx, a_gy, a_hole, a_gm, use_graph=mode)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[key] = best
res["kernel_launches"] = launches
n_eq_pix = int((np.stack([PO.equation_codes(hole[f], gm[f], np.zeros_like(hole[f])) for f in range(F)]) & 15).astype(bool).sum())
# algorithmic bytes of one iteration per pixel that owns equations (fp64, 3 channels):
#   psn_v: read u 4x3 + neighbours' u 4x3 + v 3, write v 3            = 30 doubles
#   psn_ux: read v 3 + neighbours' v 4x3 + w 3 + x 3 + u 4x3, write x, w 3+3, u 4x3 = 51 doubles
res["algorithmic_bytes_per_iteration"] = n_eq_pix * 81 * 8
iters = -(-(int(itn.max()) + 1) // P.CHUNK) * P.CHUNK      # whole chunks are replayed
res["achieved_GBps_over_all_iterations"] = res["algorithmic_bytes_per_iteration"] * iters / (res["ms_per_clip_device_resident"] * 1e-3) / 1e9
res["us_per_iteration"] = res["ms_per_clip_device_resident"] * 1e3 / iters

# CPU oracle on a bounded sample: frame 0, channel 0 (one of the 30 systems)
code = PO.equation_codes(hole[0], gm[0], np.zeros_like(hole[0]))
b = PO.rhs(code, trg[0], gx[0].astype(np.float64), gy[0].astype(np.float64))
t0 = time.perf_counter()
xo, istop_o, itn_o = PO.lsqr(code, b[..., 0])
dt = time.perf_counter() - t0
res["cpu_oracle"] = dict(sample="frame 0, channel 0 (1 of 30 systems)", seconds=dt, itn=itn_o, istop=istop_o,
                         seconds_per_clip_extrapolated=dt * 3 * F, cores=1)
res["max_abs_diff_vs_oracle_sample"] = float(np.abs(out[0, :, :, 0].cpu().numpy()[hole[0]] - xo.astype(np.float32).astype(np.float64)[hole[0]]).max())
res["itn_gpu_sample"] = int(itn[0, 0])
print(json.dumps(res))
