"""FGT forward at the other BASELINE configurations: config 4's longest window (T=18 at
432x240) and a config-5 window (720x1280, T=13), next to config 2 (432x240, T=10). CUDA-graph replay,
CUDA-event timed; per-kernel algorithmic TFLOP/s / GB/s from an instrumented eager forward.
Prints one JSON object. Run under gpurun."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, synth  # noqa: E402
from fgt_b200.fgt_model import Model  # noqa: E402

dev = torch.device("cuda:0")
cfg = dict(synth.CFG_A)
sd = synth.make_state_dict(synth.fgt_param_shapes(cfg), seed=1)
model = Model(cfg)
model.load_state_dict(sd)
model = model.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for name, (t, H, W) in {"c2_432x240_T10": (10, 240, 432), "c4_432x240_T18": (18, 240, 432),
                        "c5_1280x720_T13": (13, 720, 1280)}.items():
    clip = [x.to(dev) for x in synth.fgt_inputs(seed=2, t=t, H=H, W=W)]
    with torch.no_grad():
        model.net.enable_cuda_graph(True)
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
        model.net.enable_cuda_graph(False)
        model(*clip)
        lib.profile_start()
        model(*clip)
        recs = lib.profile_stop()
    agg = {}
    for kern, tag, fl, by, m, _sc in recs:
        key = kern if kern != "flash" else ("flash_temporal" if tag.startswith("t") else "flash_spatial")
        a = agg.setdefault(key, [0.0, 0.0, 0.0])
        a[0] += m
        a[1] += fl
        a[2] += by
    kern = {}
    for k, (m, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        kern[k] = {"ms": round(m, 3), **({"tflops": round(fl / m / 1e9, 1)} if fl else {"gbs": round(by / m / 1e6, 1)})}
    res[name] = {"ms_per_forward": round(ms, 3), "frames_per_s": round(t / ms * 1e3, 1), "kernels": kern,
                 "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
    model.net._geo.clear()  # drop this geometry's workspaces before the next one
    torch.cuda.empty_cache()
print(json.dumps(res))
