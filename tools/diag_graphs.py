"""Checks that CUDA-graph replay reproduces the eager results bit for bit and times both."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import synth  # noqa: E402


def bench(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


with torch.no_grad():
    from fgt_b200.lafc_model import Model as LAFC
    lm = LAFC(synth.CFG_LAFC)
    lm.load_state_dict(synth.make_state_dict(synth.lafc_param_shapes(), seed=5))
    lm = lm.cuda()
    fl, mk = [t.cuda() for t in synth.lafc_inputs(seed=6, H=240, W=432)]
    ref = lm(fl, mk)
    t_eager = bench(lambda: lm(fl, mk))
    lm.net.enable_cuda_graph()
    for _ in range(3):
        out = lm(fl, mk)
    print("LAFC graph == eager:", torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]),
          f"eager {t_eager:.2f} ms -> graph {bench(lambda: lm(fl, mk)):.2f} ms")

    from fgt_b200.raft_model import RAFT
    rm = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False))
    rm.load_state_dict(synth.raft_state_dict(seed=4))
    rm = rm.cuda().eval()
    a, b = [t.cuda() for t in synth.raft_inputs(seed=5, H=480, W=864)]
    ref = rm(a, b, iters=20, test_mode=True)
    t_eager = bench(lambda: rm(a, b, iters=20, test_mode=True), 5)
    rm.enable_cuda_graph()
    for _ in range(3):
        out = rm(a, b, iters=20, test_mode=True)
    print("RAFT graph == eager:", torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]),
          f"eager {t_eager:.2f} ms -> graph {bench(lambda: rm(a, b, iters=20, test_mode=True), 5):.2f} ms")

    from bench import build_model, T, H, W
    fm, _ = build_model(torch.device("cuda:0"))
    clip = [t.cuda() for t in synth.fgt_inputs(seed=3, t=T, H=H, W=W)]
    ref = fm(*clip)
    t_eager = bench(lambda: fm(*clip))
    fm.net.enable_cuda_graph()
    for _ in range(3):
        out = fm(*clip)
    print("FGT graph == eager:", torch.equal(out, ref), f"eager {t_eager:.2f} ms -> graph {bench(lambda: fm(*clip)):.2f} ms")
