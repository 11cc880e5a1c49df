"""GPU bring-up diagnostic for the fused attention kernel (dense zones and windowed + global keys)."""
import math
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import lib  # noqa: E402
from tools.diag_gemm import report  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def dense_case(batches, heads, L, qscale=1.0):
    C = heads * 128
    q = torch.randn(batches, L, C, device=dev) * qscale
    k = torch.randn(batches, L, C, device=dev)
    v = torch.randn(batches, L, C, device=dev)
    qs, ks = lib.to_split(q), lib.to_split(k)
    vs = lib.to_split(v)
    out = lib.empty_split((batches, L, C), dev)
    out.fill_(float("nan"))
    lib.attention(qs, ks, vs, out, batches=batches, heads=heads, Lq=L, Lk=L, q_ld=C, k_ld=C, v_ld=C, out_ld=C,
                  q_batch_stride=L * C, k_batch_stride=L * C, v_batch_stride=L * C, out_batch_stride=L * C,
                  scale=1 / math.sqrt(128))
    torch.cuda.synchronize()
    qh = q.double().reshape(batches, L, heads, 128).transpose(1, 2)
    kh = k.double().reshape(batches, L, heads, 128).transpose(1, 2)
    vh = v.double().reshape(batches, L, heads, 128).transpose(1, 2)
    ref = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(128), -1) @ vh
    ref = ref.transpose(1, 2).reshape(batches, L, C)
    return report(f"dense attn batches={batches} heads={heads} L={L} qscale={qscale}", lib.from_split(out), ref)


def window_case(frames, heads, nwin, nglob, qscale=1.0):
    C = heads * 128
    nwp = (nwin + 1) // 2 * 2
    gl_rows = (nglob + 63) // 64 * 64
    rows = nwp * 64 + gl_rows
    q = torch.randn(frames, nwp * 64, C, device=dev) * qscale
    k = torch.randn(frames, rows, C, device=dev)
    v = torch.randn(frames, rows, C, device=dev)
    qs, ks = lib.to_split(q), lib.to_split(k)
    vs = lib.to_split(v)
    out = lib.empty_split((frames, nwp * 64, C), dev)
    out.fill_(float("nan"))
    lib.attention(qs, ks, vs, out, batches=frames, heads=heads, Lq=nwp * 64, Lk=rows, Lk_rows=rows, q_ld=C, k_ld=C,
                  v_ld=C, out_ld=C, q_batch_stride=nwp * 64 * C, k_batch_stride=rows * C,
                  v_batch_stride=rows * C, out_batch_stride=nwp * 64 * C, scale=1 / math.sqrt(128), mode=1,
                  glob_start=nwp * 64, glob_count=nglob)
    torch.cuda.synchronize()
    ref = torch.empty(frames, nwp * 64, C, dtype=torch.double, device=dev)
    for w in range(nwp):
        qw = q[:, w * 64:(w + 1) * 64].double().reshape(frames, 64, heads, 128).transpose(1, 2)
        kk = torch.cat([k[:, w * 64:(w + 1) * 64], k[:, nwp * 64:nwp * 64 + nglob]], 1).double()
        vv = torch.cat([v[:, w * 64:(w + 1) * 64], v[:, nwp * 64:nwp * 64 + nglob]], 1).double()
        kh = kk.reshape(frames, -1, heads, 128).transpose(1, 2)
        vh = vv.reshape(frames, -1, heads, 128).transpose(1, 2)
        o = torch.softmax(qw @ kh.transpose(-1, -2) / math.sqrt(128), -1) @ vh
        ref[:, w * 64:(w + 1) * 64] = o.transpose(1, 2).reshape(frames, 64, C)
    return report(f"window attn frames={frames} heads={heads} nwin={nwin} nglob={nglob}", lib.from_split(out), ref)


def main():
    cases = [
        lambda: dense_case(1, 1, 64),
        lambda: dense_case(1, 1, 128),
        lambda: dense_case(1, 2, 200),
        lambda: dense_case(2, 4, 1800, qscale=3.0),
        lambda: dense_case(4, 4, 37),
        lambda: window_case(1, 1, 2, 60),
        lambda: window_case(3, 4, 15, 60, qscale=3.0),
        lambda: window_case(1, 4, 112, 448),
    ]
    nfail = 0
    for c in cases:
        try:
            if not c():
                nfail += 1
        except Exception:
            traceback.print_exc()
            nfail += 1
            break
    print("FAILURES:", nfail, flush=True)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
