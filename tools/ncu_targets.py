"""Small launch sets for ncu captures: `python tools/ncu_targets.py gemm|flash|model` runs the named
kernels a few times on representative FGT shapes (432x240, T=10)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, packing, synth  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else "gemm"


def linear(M, N, K, bn, reps=3):
    a = lib.to_split(torch.randn(M, K, device=dev))
    w = packing.pack_weight(torch.randn(N, K, device=dev) / K ** 0.5).to(dev)
    b = torch.randn(N, device=dev)
    out = lib.empty_split((M, N), dev)
    for _ in range(reps):
        lib.gemm_tc([lib.ASeg(a, K, M)], w, N, out_w=M, bn=bn, bias=b, out_split=out, tag=f"lin{M}x{N}x{K}")
    torch.cuda.synchronize()


def conv(n, h, w_, cin, cout, bn, reps=3):
    x = lib.to_split(torch.randn(n, h, w_, cin, device=dev))
    wt = packing.pack_weight(torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5).to(dev)
    b = torch.randn(cout, device=dev)
    out = lib.empty_split((n, h, w_, cout), dev)
    for _ in range(reps):
        lib.gemm_tc([lib.ASeg(x, cin, w_, h, n)], wt, cout, kx=3, ky=3, pad_x=1, pad_y=1, out_w=w_, out_h=h, out_z=n,
                    box_w=16, box_h=8, bn=bn, bias=b, act=lib.ACT_LEAKY02, out_split=out, os_z=h * w_ * cout,
                    os_y=w_ * cout, os_x=cout, tag=f"conv{cin}->{cout}")
    torch.cuda.synchronize()


if what == "gemm":
    linear(7200, 1024, 512, 128)      # temporal QK projection
    linear(7200, 1960, 512, 128)      # fusion FFN conv1
    linear(7200, 512, 1960, 128)      # fusion FFN conv2
    conv(10, 60, 108, 256, 384, 128)  # encoder layer 8
    conv(10, 240, 432, 64, 64, 64)    # decoder layer 3
elif what == "flash":
    from tools import diag_attn as D
    D.dense_case(4, 4, 1800, qscale=3.0)
    D.window_case(10, 4, 15, 60, qscale=3.0)
else:
    from bench import build_model, T, H, W
    model, _ = build_model(dev)
    clip = [t.to(dev) for t in synth.fgt_inputs(seed=3, t=T, H=H, W=W)]
    with torch.no_grad():
        for _ in range(2):
            model(*clip)
    torch.cuda.synchronize()
