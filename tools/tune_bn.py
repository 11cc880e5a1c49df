"""Times representative linear shapes for different output-channel tiles (bn)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, packing  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def t_linear(M, N, K, bn, split_out=True, reps=20):
    a = lib.to_split(torch.randn(M, K, device=dev))
    w = packing.pack_weight(torch.randn(N, K, device=dev) / K ** 0.5).to(dev)
    b = torch.randn(N, device=dev)
    out = lib.empty_split((M, N), dev) if split_out else torch.empty(M, N, device=dev)
    kw = dict(out_split=out) if split_out else dict(out_f32=out)
    f = lambda: lib.gemm_tc([lib.ASeg(a, K, M)], w, N, out_w=M, bn=bn, bias=b, **kw)  # noqa: E731
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * M * N * K / ms / 1e9


for (M, N, K, so) in [(7200, 1024, 512, True), (7200, 512, 512, True), (7200, 1960, 512, False), (7200, 512, 1960, False),
                      (10880, 1024, 768, True), (7200, 6272, 512, False)]:
    row = []
    for bn in (64, 96, 128, 160, 192, 256):
        if bn > N:
            continue
        ms, tf = t_linear(M, N, K, bn, so)
        row.append(f"bn{bn}: {ms * 1e3:6.1f}us {tf:5.0f}TF")
    print(f"M={M} N={N} K={K}: " + " | ".join(row), flush=True)
