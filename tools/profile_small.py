"""Forward time of short windows (the per-rank problem of a frame-sharded window): CUDA-graph replay ms at
T = 1, 2, 3, 4, 6, 10 frames of 432x240, and the per-launch breakdown at T = 2. Run with FGT_PDL=0 / 1."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, synth  # noqa: E402
from bench import build_model  # noqa: E402

dev = torch.device("cuda:0")
model, _ = build_model(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"pdl": os.environ.get("FGT_PDL", "0")}
for t in (1, 2, 3, 4, 6, 10):
    clip = [x.to(dev) for x in synth.fgt_inputs(seed=3, t=t, H=240, W=432)]
    model.net.enable_cuda_graph(True)
    with torch.no_grad():
        for _ in range(4):
            model(*clip)
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            model(*clip)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b) / 10
    res[f"T{t}_ms"] = round(ms, 3)
print(json.dumps(res))
if "--layers" in sys.argv:
    model.net.enable_cuda_graph(False)
    clip = [x.to(dev) for x in synth.fgt_inputs(seed=3, t=2, H=240, W=432)]
    with torch.no_grad():
        model(*clip)
        lib.profile_start()
        for _ in range(3):
            model(*clip)
        recs = lib.profile_stop()
    n = len(recs) // 3
    tot = 0.0
    for i in range(n):
        ms = sum(recs[i + r * n][4] for r in range(3)) / 3
        tot += ms
        print(f"{recs[i][0]:14s} {recs[i][1]:14s} {ms * 1e3:8.1f} us")
    print(f"sum of launches at T=2: {tot:.3f} ms over {n} launches")
