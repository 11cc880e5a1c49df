"""Per-launch timing (CUDA events) of one FGT forward at 432x240 T=10: prints every launch with its
tag, milliseconds, algorithmic TFLOP/s or GB/s. Run under gpurun."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, synth  # noqa: E402
from bench import build_model, T, H, W  # noqa: E402

dev = torch.device("cuda:0")
model, sd = build_model(dev)
clip = [t.to(dev) for t in synth.fgt_inputs(seed=3, t=T, H=H, W=W)]
with torch.no_grad():
    for _ in range(3):
        model(*clip)
    reps = 5
    lib.profile_start()
    for _ in range(reps):
        model(*clip)
    recs = lib.profile_stop()
n = len(recs) // reps
tot = 0.0
print(f"{'kernel':14s} {'tag':14s} {'ms':>8s} {'TFLOP/s':>9s} {'GB/s':>8s}")
for i in range(n):
    ms = sum(recs[i + r * n][4] for r in range(reps)) / reps
    k, tag, fl, by, _, _sc = recs[i]
    tot += ms
    tf = fl / (ms * 1e-3) / 1e12 if fl else 0.0
    gb = by / (ms * 1e-3) / 1e9
    print(f"{k:14s} {tag:14s} {ms:8.4f} {tf:9.1f} {gb:8.0f}")
print(f"sum of launches: {tot:.3f} ms")
