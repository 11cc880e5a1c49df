"""CPU emulation of tensor-core operand schemes per layer class.

Runs the CPU oracle (oracle/fgt_oracle.py) with the operands of every contraction of one layer class
rounded the way a given MMA scheme would see them, fp32 accumulation, and reports the end-to-end error of
the FGT forward against the un-rounded fp32 run, for both weight regimes. Emulated schemes:

    bf16x3  hi/lo split-bf16 operands, 3 MMAs (hi*hi + hi*lo + lo*hi)            cost 3
    f16x2a  activation split into fp16 hi+lo, weight rounded to fp16, 2 MMAs      cost 2
    f16x2w  weight split, activation rounded to fp16, 2 MMAs                      cost 2
    f16     both operands rounded to fp16, 1 MMA                                  cost 1
    tf32    both operands rounded (RN) to tf32, 1 kind::tf32 MMA (half rate)      cost 2
    bf16    both operands rounded to bf16, 1 MMA                                  cost 1

Test infrastructure (imports oracle/); never part of the product path.

    python tools/precision_study.py [--t 4] [--mode loo|uniform|mix] [--mix cls=scheme,...]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fgt_b200 import synth  # noqa: E402
from oracle import fgt_oracle as O  # noqa: E402

COST = {"f32": 0, "bf16x3": 3, "f16x2a": 2, "f16x2w": 2, "f16": 1, "tf32": 2, "bf16": 1}


def q16(x):
    y = x.half()
    if torch.isinf(y).any():
        raise OverflowError("fp16 overflow")
    return y.float()


def qb(x):
    return x.bfloat16().float()


def qsplit_b(x):
    h = qb(x)
    return h + qb(x - h)


def qtf32(x):
    i = x.contiguous().view(torch.int32)
    r = ((i + 0x0FFF + ((i >> 13) & 1)) >> 13) << 13  # round to nearest even on 13 dropped bits
    return r.view(torch.float32)


def quant(a, w, scheme):
    if scheme == "f32":
        return a, w
    if scheme == "bf16x3":
        return qsplit_b(a), qsplit_b(w)
    if scheme == "f16x2a":
        return a, q16(w)
    if scheme == "f16x2w":
        return q16(a), w
    if scheme == "f16":
        return q16(a), q16(w)
    if scheme == "tf32":
        return qtf32(a), qtf32(w)
    if scheme == "bf16":
        return qb(a), qb(w)
    raise ValueError(scheme)


CLASSES = ["enc", "fenc", "p2v", "fp2v", "tproj", "tout", "sgate", "sproj", "sout", "tqk", "tpv", "sqk", "spv", "ffn1", "ffn2",
           "v2p", "dec"]


def classify(key):
    if key.startswith("frame_endoder"):
        return "enc"
    if key.startswith("flow_encoder"):
        return "fenc"
    if key.startswith("patch2vec"):
        return "p2v"
    if key.startswith("f_patch2vec"):
        return "fp2v"
    if key.startswith("vec2patch"):
        return "v2p"
    if key.startswith("decoder"):
        return "dec"
    if "ffn.conv1" in key:
        return "ffn1"
    if "ffn.conv2" in key:
        return "ffn2"
    spatial = "s_transformer" in key
    if "reweightFlow" in key:
        return "sgate"
    if "output_linear" in key:
        return "sout" if spatial else "tout"
    if "_embedding" in key:
        return "sproj" if spatial else "tproj"
    return None  # depthwise convs, norms: CUDA-core fp32


class Emu:
    """Patches the oracle's F.conv2d / F.linear / sdpa with operand-rounded versions."""

    def __init__(self, sd, mix):
        self.mix = mix
        self.cls = {}
        for k, v in sd.items():
            if k.endswith(".weight") and v.dim() >= 2:
                c = classify(k)
                if c:
                    self.cls[id(v)] = c
        self.F = torch.nn.functional

    def scheme(self, w):
        return self.mix.get(self.cls.get(id(w)), "f32")

    def conv2d(self, x, w, b=None, **kw):
        a, ww = quant(x, w, self.scheme(w))
        return self.F.conv2d(a, ww, b, **kw)

    def linear(self, x, w, b=None):
        a, ww = quant(x, w, self.scheme(w))
        return self.F.linear(a, ww, b)

    def sdpa(self, q, k, v):
        import math
        t = q.dim() == 4
        a, b = quant(q, k, self.mix.get("tqk" if t else "sqk", "f32"))
        s = a @ b.transpose(-2, -1) / math.sqrt(q.shape[-1])
        # the kernel feeds un-normalised exponentials (<= 2^8) to the PV MMA and divides at the end
        m = s.amax(-1, keepdim=True)
        p = torch.exp(s - m)
        pp, vv = quant(p, v, self.mix.get("tpv" if t else "spv", "f32"))
        return (pp @ vv) / p.sum(-1, keepdim=True)

    def __getattr__(self, name):  # everything else straight from torch.nn.functional
        return getattr(self.F, name)


def run(sd, clip, mix):
    emu = Emu(sd, mix)
    old_F, old_sdpa = O.F, O.sdpa
    O.F, O.sdpa = emu, emu.sdpa
    try:
        with torch.no_grad():
            return O.fgt_forward(sd, *clip)
    finally:
        O.F, O.sdpa = old_F, old_sdpa


def err(out, ref):
    d = (out.double() - ref.double())
    return (d.norm() / ref.double().norm()).item(), (d.abs().max() / ref.double().abs().max()).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--t", type=int, default=4)
    ap.add_argument("--H", type=int, default=240)
    ap.add_argument("--W", type=int, default=432)
    ap.add_argument("--mode", default="loo", choices=["loo", "uniform", "mix"])
    ap.add_argument("--base", default="bf16x3")
    ap.add_argument("--demote", default="f16")
    ap.add_argument("--mix", default="")
    ap.add_argument("--regimes", default="scaled,default")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    res = []
    for regime in a.regimes.split(","):
        cfg = dict(synth.CFG_A)
        cfg["input_resolution"] = (a.H, a.W)
        sd = O.strip_net(synth.make_state_dict(synth.fgt_param_shapes(cfg), seed=1, regime=regime))
        clip = synth.fgt_inputs(seed=3, t=a.t, H=a.H, W=a.W)
        ref = run(sd, clip, {})
        jobs = []
        if a.mode == "uniform":
            for s in ("bf16x3", "f16x2a", "f16x2w", "f16", "tf32", "bf16"):
                jobs.append((f"all={s}", {c: s for c in CLASSES}))
        elif a.mode == "loo":
            jobs.append((f"all={a.base}", {c: a.base for c in CLASSES}))
            for c in CLASSES:
                m = {k: a.base for k in CLASSES}
                m[c] = a.demote
                jobs.append((f"{c}={a.demote}", m))
        else:
            m = {c: a.base for c in CLASSES}
            for kv in a.mix.split(","):
                if kv:
                    k, v = kv.split("=")
                    m[k] = v
            jobs.append((a.mix or "base", m))
        for name, m in jobs:
            t0 = time.time()
            try:
                e = err(run(sd, clip, m), ref)
            except OverflowError as ex:
                e = (float("nan"), float("nan"))
                name += f" [{ex}]"
            r = dict(regime=regime, job=name, rel_l2=e[0], max_over_max=e[1], sec=round(time.time() - t0, 1))
            res.append(r)
            print(json.dumps(r), flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
