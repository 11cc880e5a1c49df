"""Small launch sets for ncu captures: `python tools/ncu_targets.py <what>` runs the named kernels a few times on
representative shapes (FGT 432x240 T=10; RAFT 480x864; LAFC 240x432; propagation / region fill / Poisson 240x432x10).

    gemm | flash | tail | model | raft | lafc | prop | fill | poisson | splat
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, packing, synth  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else "gemm"


def linear(M, N, K, bn, reps=3, f32_out=False):
    a = lib.to_split(torch.randn(M, K, device=dev))
    w = packing.pack_weight(torch.randn(N, K, device=dev) / K ** 0.5).to(dev)
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev) if f32_out else lib.empty_split((M, N), dev)
    for _ in range(reps):
        lib.gemm_tc([lib.ASeg(a, K, M)], w, N, out_w=M, bn=bn, bias=b, tag=f"lin{M}x{N}x{K}",
                    **({"out_f32": out} if f32_out else {"out_split": out}))
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
    linear(7200, 1536, 512, 128)      # temporal Q|K|V projection
    linear(7200, 1960, 512, 128, f32_out=True)  # fusion FFN conv1
    linear(7200, 512, 1960, 128)      # fusion FFN conv2
    conv(10, 60, 108, 256, 384, 128)  # encoder layer 8
    conv(10, 240, 432, 64, 64, 64)    # decoder layer 3
elif what == "flash":
    from tools import diag_attn as D
    D.dense_case(4, 4, 1800, qscale=3.0)
    D.window_case(10, 4, 15, 60, qscale=3.0)
elif what == "tail":
    n, H, W, cin = 10, 240, 432, 64
    xs = lib.to_split(torch.randn(n, H, W, cin, device=dev))
    w = torch.randn(3, cin, 3, 3, device=dev) / 24
    ws = packing.pack_weight(lib.pack_taps_as_n(w)).to(dev)
    out = torch.empty(n, 3, H, W, device=dev)
    for _ in range(3):
        lib.conv_tail(xs, n, H, W, cin, ws, 3, torch.zeros(3, device=dev), lib.ACT_TANH, out, nchw=True)
    torch.cuda.synchronize()
elif what == "model":
    from bench import build_model, T, H, W
    model, _ = build_model(dev)
    clip = [t.to(dev) for t in synth.fgt_inputs(seed=3, t=T, H=H, W=W)]
    with torch.no_grad():
        for _ in range(2):
            model(*clip)
    torch.cuda.synchronize()
elif what == "raft":
    from fgt_b200.raft_model import RAFT
    m = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False))
    m.load_state_dict(synth.raft_state_dict(seed=4))
    m = m.cuda().eval()
    im1, im2 = synth.raft_inputs(seed=5, H=480, W=864)
    with torch.no_grad():
        for _ in range(2):
            m(im1.cuda(), im2.cuda(), iters=4, test_mode=True)   # 4 refinement iterations: every kernel of the loop appears
    torch.cuda.synchronize()
elif what == "lafc":
    from fgt_b200.lafc_model import Model as LAFC
    lm = LAFC(synth.CFG_LAFC)
    lm.load_state_dict(synth.make_state_dict(synth.lafc_param_shapes(), seed=5))
    lm = lm.cuda()
    fl, mk = synth.lafc_inputs(seed=6, H=240, W=432)
    with torch.no_grad():
        for _ in range(2):
            lm(fl.cuda(), mk.cuda())
    torch.cuda.synchronize()
elif what == "prop":
    from fgt_b200.propagation import get_flowNN_gradient
    gx, gy, mask, ff, fb = synth.prop_inputs(seed=7, H=240, W=432, N=10)
    args = argparse.Namespace(Nonlocal=False, consistencyThres=5.0, alpha=0.1)
    for _ in range(2):
        get_flowNN_gradient(args, gx.copy(), gy.copy(), mask, mask, ff, fb)
elif what == "fill":
    from fgt_b200 import regionfill as RF
    img, mask = synth.regionfill_inputs(seed=8, B=18, H=240, W=432)
    mask[-1, 60:140, 100:260] = True
    img[-1][mask[-1]] = 0
    flows = np.stack([img, img[::-1].copy()], -1)
    RF.diffusion(flows, mask[..., None])
elif what == "poisson":
    from fgt_b200.poisson import poisson_blend_batch
    trg, gx, gy, hole, gm = synth.poisson_inputs(seed=9, F=10, H=240, W=432)[:5]
    poisson_blend_batch(trg, gx, gy, hole, gm, use_graph=False)
    torch.cuda.synchronize()
elif what == "splat":
    from fgt_b200 import flow_warp as FW
    feat, flow = synth.flow_warp_inputs(seed=1, b=2, c=64, h=240, w=432)
    for _ in range(2):
        FW.flow_prop(feat.cuda(), flow.cuda(), "forward")
    torch.cuda.synchronize()
else:
    raise SystemExit(f"unknown target {what}")
