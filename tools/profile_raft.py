"""Per-launch timing of one RAFT forward (480x864, 20 iterations): aggregates by tag."""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, synth  # noqa: E402
from fgt_b200.raft_model import RAFT  # noqa: E402

m = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False))
m.load_state_dict(synth.raft_state_dict(seed=4))
m = m.cuda().eval()
a, b = [t.cuda() for t in synth.raft_inputs(seed=5, H=480, W=864)]
with torch.no_grad():
    for _ in range(2):
        m(a, b, iters=20, test_mode=True)
    lib.profile_start()
    m(a, b, iters=20, test_mode=True)
    recs = lib.profile_stop()
agg = collections.OrderedDict()
for k, tag, fl, by, ms, _sc in recs:
    key = f"{k}:{tag}"
    e = agg.setdefault(key, [0, 0.0, 0.0])
    e[0] += 1
    e[1] += ms
    e[2] += fl
tot = sum(e[1] for e in agg.values())
for key, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{key:34s} n={n:3d} total={ms:7.3f} ms  per={ms / n * 1e3:7.1f} us  {fl / ms / 1e9 if fl else 0:6.1f} TF")
print(f"sum {tot:.3f} ms over {len(recs)} launches")
