"""Per-key-tile timeline (clock64) of one CTA of the temporal flash-attention launch."""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib  # noqa: E402

dev = torch.device("cuda:0")
L, C, heads, batches = 1800, 512, 4, 4
q = lib.to_split(torch.randn(batches, L, C, device=dev) * 3)
k = lib.to_split(torch.randn(batches, L, C, device=dev))
v = lib.to_split(torch.randn(batches, L, C, device=dev))
out = lib.empty_split((batches, L, C), dev)
trace = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
L_ = lib.load()
L_.fgt_debug_flash_trace.argtypes = [ctypes.c_void_p]


def run():
    lib.attention(q, k, v, out, batches=batches, heads=heads, Lq=L, Lk=L, q_ld=C, k_ld=C, v_ld=C, out_ld=C,
                  q_batch_stride=L * C, k_batch_stride=L * C, v_batch_stride=L * C, out_batch_stride=L * C,
                  scale=1 / math.sqrt(128))


for _ in range(3):
    run()
L_.fgt_debug_flash_trace(ctypes.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
L_.fgt_debug_flash_trace(None)
t = trace.cpu().reshape(3, 64, 8)
t0 = int(t[1, 0, 0])
print("tile | MMA: kfull sempty S_issued vfull pfull PV_issued | SOFTMAX: start sfull ld_done compute_done pempty arrived | PROD: kempty vempty")
for j in range(8, 16):
    m = [int(x) - t0 for x in t[1, j, :6]]
    s = [int(x) - t0 for x in t[2, j, :6]]
    pr = [int(x) - t0 for x in t[0, j, :2]]
    print(j, "|", " ".join(f"{x:7d}" for x in m), "|", " ".join(f"{x:7d}" for x in s), "|", " ".join(f"{x:7d}" for x in pr))
per = (int(t[2, 20, 5]) - int(t[2, 8, 5])) / 12
print("cycles per tile (softmax arrive to arrive):", per)
