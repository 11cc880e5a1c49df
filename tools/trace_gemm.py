"""clock64 timeline of one CTA of gemm_tc_kernel on a transformer linear (fgt_debug_gemm_trace): per local tile, when
the producer, the MMA issuer and the epilogue started, what they waited for and when they finished — to see whether a
K=512 linear loses its time to the wave tail, to operands arriving late, to a busy accumulator buffer or to the
epilogue. Run under gpurun:  python tools/trace_gemm.py [M N K bn]   (default: temporal QK projection 7200 1024 512 128)

Slots (cycles relative to the CTA's first producer event):
  PROD start | slot_free (first stage of the tile could be overwritten) | issued (last load of the tile issued)
  MMA  start | acc_free (accumulator buffer released by the epilogue) | ops_in (first operands landed) | committed
  EPI  start | acc_full (all MMAs of the tile retired) | stored
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib, packing  # noqa: E402

dev = torch.device("cuda:0")
M, N, K, bn = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (7200, 1024, 512, 128)
torch.manual_seed(0)
a = lib.to_split(torch.randn(M, K, device=dev))
w = packing.pack_weight(torch.randn(N, K, device=dev) / K ** 0.5).to(dev)
b = torch.randn(N, device=dev)
out = lib.empty_split((M, N), dev)
L = lib.load()
L.fgt_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.fgt_debug_gemm_trace.restype = ctypes.c_int


def run():
    lib.gemm_tc([lib.ASeg(a, K, M)], w, N, out_w=M, bn=bn, bias=b, out_split=out, tag="trace")


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
tiles = -(-M // 128) * -(-N // bn)
print(f"linear {M}x{N}x{K} bn={bn}: {us:.1f} us per launch (untraced), {tiles} tiles on 148 SMs "
      f"({tiles / 148:.2f} waves), {2 * M * N * K / us / 1e6:.0f} TFLOP/s algorithmic")
for cta in (0, 147 if tiles > 147 else tiles - 1):
    trace = torch.zeros(3 * 64 * 4, dtype=torch.int64, device=dev)
    lib.check(L.fgt_debug_gemm_trace(ctypes.c_void_p(trace.data_ptr()), cta), "fgt_debug_gemm_trace")
    run()
    torch.cuda.synchronize()
    lib.check(L.fgt_debug_gemm_trace(None, 0), "fgt_debug_gemm_trace")
    t = trace.cpu().reshape(3, 64, 4)
    n_local = int((t[1, :, 3] != 0).sum())
    t0 = int(t[0, 0, 0])
    print(f"CTA {cta}: {n_local} tiles")
    print("tile | PROD start slot_free   issued | MMA start acc_free   ops_in committed | EPI start acc_full   stored")
    for j in range(n_local):
        pr = [int(x) - t0 for x in t[0, j, :3]]
        mm = [int(x) - t0 for x in t[1, j, :4]]
        ep = [int(x) - t0 for x in t[2, j, :3]]
        print(f"{j:4d} | " + " ".join(f"{x:9d}" for x in pr) + " | " + " ".join(f"{x:8d}" for x in mm) + " | " +
              " ".join(f"{x:8d}" for x in ep))
    if n_local:
        span = int(t[2, n_local - 1, 2]) - t0
        mma_busy = sum(int(t[1, j, 3]) - int(t[1, j, 2]) for j in range(n_local))
        print(f"  CTA span {span} cycles; MMA issue windows {mma_busy} cycles ({mma_busy / span:.0%}); "
              f"epilogue of the last tile {int(t[2, n_local - 1, 2]) - int(t[2, n_local - 1, 1])} cycles after its MMAs retired")
