"""Stage timings for the rows next to the FGT transformer (BASELINE config 3 shapes): RAFT pair at
480x864 (20 iterations), LAFC call at 240x432 (3 candidate flows), get_flowNN_gradient on a
240x432x10 clip — CUDA-event timed, with the per-kernel breakdown (algorithmic FLOPs / bytes) and
the CPU oracle next to it. Prints one JSON object. Run under gpurun."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, synth  # noqa: E402
from bench import load_peaks  # noqa: E402

dev = torch.device("cuda:0")
peaks = load_peaks()
quick = "--quick" in sys.argv


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / reps


def breakdown(fn, reps=3):
    lib.profile_start()
    for _ in range(reps):
        fn()
    recs = lib.profile_stop()
    agg = {}
    for k, tag, fl, by, ms, _sc in recs:
        a = agg.setdefault(k, dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
        a["ms"] += ms / reps
        a["flops"] += fl / reps
        a["bytes"] += by / reps
        a["n"] += 1 / reps
    out = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        e = dict(ms=round(a["ms"], 4), launches=round(a["n"]))
        if a["flops"]:
            e["tflops"] = round(a["flops"] / a["ms"] / 1e9, 1)
            e["frac_of_bf16_peak"] = round(e["tflops"] / peaks["tf_sustained"], 4)
        else:
            e["gbs"] = round(a["bytes"] / a["ms"] / 1e6, 1)
            e["frac_of_hbm_peak"] = round(e["gbs"] / peaks["hbm_gbs"], 4)
        out[k] = e
    return out


res = {}
with torch.no_grad():
    # ---- RAFT
    from fgt_b200.raft_model import RAFT
    from oracle import raft_oracle as RO
    sd = synth.raft_state_dict(seed=4)
    m = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False))
    m.load_state_dict(sd)
    m = m.cuda().eval()
    im1, im2 = synth.raft_inputs(seed=5, H=480, W=864)
    a, b = im1.cuda(), im2.cuda()
    f = lambda: m(a, b, iters=20, test_mode=True)  # noqa: E731
    ms = timed(f)
    r = dict(ms_per_pair=ms, pairs_per_s=1e3 / ms, kernels=breakdown(f))
    if not quick:
        t0 = time.perf_counter()
        RO.raft_forward(sd, im1, im2, iters=20)
        r["cpu_oracle_s"] = time.perf_counter() - t0
    res["raft_480x864_i20"] = r
    # batched + CUDA-graph replay: what a clip driver should call (forward and backward flows of several
    # frame pairs per call)
    m.enable_cuda_graph(True)
    for nb in (1, 4):
        ib1, ib2 = synth.raft_inputs(seed=5, H=480, W=864, n=nb)
        ga, gb = ib1.cuda(), ib2.cuda()
        ms = timed(lambda: m(ga, gb, iters=20, test_mode=True))
        res[f"raft_480x864_i20_graph_batch{nb}"] = dict(ms_per_call=ms, ms_per_pair=ms / nb, pairs_per_s=1e3 * nb / ms)
    m.enable_cuda_graph(False)
    # ---- LAFC
    from fgt_b200.lafc_model import Model as LAFC
    from oracle import lafc_oracle as LO
    sd = synth.make_state_dict(synth.lafc_param_shapes(), seed=5)
    lm = LAFC(synth.CFG_LAFC)
    lm.load_state_dict(sd)
    lm = lm.cuda()
    fl, mk = synth.lafc_inputs(seed=6, H=240, W=432)
    a2, b2 = fl.cuda(), mk.cuda()
    f = lambda: lm(a2, b2)  # noqa: E731
    ms = timed(f)
    r = dict(ms_per_call=ms, calls_per_s=1e3 / ms, kernels=breakdown(f))
    if not quick:
        t0 = time.perf_counter()
        LO.lafc_forward({k[4:]: v for k, v in sd.items()}, fl, mk)
        r["cpu_oracle_s"] = time.perf_counter() - t0
    res["lafc_240x432"] = r
# ---- propagation (host arrays in, host arrays out: includes H2D/D2H like the reference API)
from fgt_b200.propagation import get_flowNN_gradient  # noqa: E402
from oracle import prop_oracle as PO  # noqa: E402
gx, gy, mask, ff, fb = synth.prop_inputs(seed=7, H=240, W=432, N=10)
args = argparse.Namespace(Nonlocal=False, consistencyThres=5.0, alpha=0.1)
for _ in range(2):
    get_flowNN_gradient(args, gx.copy(), gy.copy(), mask, mask, ff, fb)
t0 = time.perf_counter()
for _ in range(5):
    get_flowNN_gradient(args, gx.copy(), gy.copy(), mask, mask, ff, fb)
r = dict(ms_per_clip_host_to_host=(time.perf_counter() - t0) / 5 * 1e3)
t0 = time.perf_counter()
PO.get_flownn_gradient(gx, gy, mask, ff, fb, 5.0, 0.1)
r["cpu_oracle_s"] = time.perf_counter() - t0
res["prop_240x432x10"] = r
# ---- flow diffusion (region fill) of the 2*(N-1) = 18 incomplete flows of a 10-frame 240x432 clip: one batch
from fgt_b200 import regionfill as RF  # noqa: E402
from oracle import regionfill_oracle as RFO  # noqa: E402
img, mask = synth.regionfill_inputs(seed=8, B=18, H=240, W=432)
mask[-1, 60:140, 100:260] = True          # the synthetic set leaves its last mask empty; use a plain box there
img[-1][mask[-1]] = 0
flows = np.stack([img, img[::-1].copy()], -1)
for _ in range(2):
    RF.diffusion(flows, mask[..., None])
t0 = time.perf_counter()
for _ in range(3):
    got = RF.diffusion(flows, mask[..., None])
r = dict(ms_per_clip_host_to_host=(time.perf_counter() - t0) / 3 * 1e3, solves=36,
         cg_iterations=RF.regionfill_batch(img, mask, return_iters=True)[1], hole_fraction=float(mask.mean()))
t0 = time.perf_counter()
ref = RFO.diffusion(flows[:2], mask[:2, ..., None])
r["cpu_oracle_s_per_clip"] = (time.perf_counter() - t0) * 9
r["max_abs_diff_vs_oracle"] = float(max(np.abs(a - b).max() for a, b in zip(got[:2], ref)))
res["diffusion_240x432_18flows"] = r
print(json.dumps(res))
