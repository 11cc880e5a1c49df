"""GPU bring-up diagnostic for the full FGT forward: compares every captured intermediate and the
final output with the CPU oracle (small geometry by default). Run under gpurun."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import synth  # noqa: E402
from fgt_b200.fgt_model import Model  # noqa: E402
from oracle import fgt_oracle as O  # noqa: E402


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def run(H, W, t, regime, res=None, seed=1):
    cfg = dict(synth.CFG_A)
    cfg['input_resolution'] = res if res is not None else (H, W)
    sd = synth.make_state_dict(synth.fgt_param_shapes(cfg), seed=seed, regime=regime)
    model = Model(cfg)
    model.load_state_dict(sd)
    model = model.cuda()
    fr, fl, mk = synth.fgt_inputs(seed=3, t=t, H=H, W=W)
    cap = {}
    model.net.capture = cap
    with torch.no_grad():
        out = model(fr.cuda(), fl.cuda(), mk.cuda())
        torch.cuda.synchronize()
        t0 = time.time()
        ref, inter = O.fgt_forward(O.strip_net(sd), fr, fl, mk, return_intermediates=True)
        t_cpu = time.time() - t0
    print(f"== {H}x{W} t={t} regime={regime} (oracle {t_cpu:.2f}s)")
    bt = t
    ok = True
    for name in ("enc", "tok0", "ftok", "t0", "s0", "tok_final"):
        got = cap[name]
        r = inter[name]
        if name == "enc":
            got = got.permute(0, 3, 1, 2)
        else:
            got = got.reshape(bt, -1, got.shape[-1])
        e = relerr(got, r)
        print(f"   {name:10s} rel_l2={e[0]:.3e} max/max={e[1]:.3e}")
    e = relerr(out, ref)
    ok = e[0] < 1e-3 and e[1] < 1e-3 and bool(torch.isfinite(out).all())
    print(f"   {'OUTPUT':10s} rel_l2={e[0]:.3e} max/max={e[1]:.3e}  -> {'OK' if ok else 'FAIL'}", flush=True)
    return ok


if __name__ == "__main__":
    okall = True
    okall &= run(64, 96, 3, "scaled")
    okall &= run(64, 96, 3, "default")
    okall &= run(72, 100, 2, "scaled", res=(64, 96))
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        okall &= run(240, 432, 10, "scaled")
    print("ALL OK" if okall else "SOME FAILED")
    sys.exit(0 if okall else 1)
