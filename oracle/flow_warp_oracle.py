"""CPU oracle for forward flow splatting (TEST INFRASTRUCTURE — see oracle/fgt_oracle.py).

Restates /root/reference/LAFC/models/utils/flow_warp.py:4-94 (`flow_prop` -> `warp` -> `sample_one` x 4 +
`get_gaussian_weights`) with one index_add per neighbour instead of the flattened put_(accumulate=True) with
materialised index tensors. Flow channel 0 shifts the column index and channel 1 the row index (:24-25: `y = flow[:, 0]`,
`x = flow[:, 1]`; :60-61: `basex` is the row arange, `basey` the column arange; :72-77: `idxx*w + idxy`). Pinned by
tests/golden/flow_warp.npz (outputs of the unmodified reference)."""
import torch


def flow_prop(feat, flow, mode="forward"):
    assert mode in ("forward", "backward")
    b, c, h, w = feat.shape
    y, x = flow[:, 0], flow[:, 1]                                  # [b,h,w]
    x1, y1 = torch.floor(x), torch.floor(y)
    rows = torch.arange(h).view(1, h, 1).expand(b, h, w)
    cols = torch.arange(w).view(1, 1, w).expand(b, h, w)
    acc = torch.zeros(b, c, h * w)
    osum = torch.zeros(b, h * w)
    sign = 1 if mode == "forward" else -1
    for xs, ys in ((x1, y1), (x1, y1 + 1), (x1 + 1, y1), (x1 + 1, y1 + 1)):      # :37-40
        wgt = torch.exp(-((x - xs) ** 2 + (y - ys) ** 2))                          # :88-93, sigma = 1
        r = rows + sign * xs.long()
        q = cols + sign * ys.long()
        ok = (r >= 0) & (r < h) & (q >= 0) & (q < w)                               # :66
        for bi in range(b):
            idx = (r[bi] * w + q[bi])[ok[bi]]
            acc[bi].index_add_(1, idx, (feat[bi] * wgt[bi])[:, ok[bi]])
            osum[bi].index_add_(0, idx, wgt[bi][ok[bi]])
    o = osum.view(b, 1, h, w).expand(b, c, h, w)
    out = acc.view(b, c, h, w).clone()
    out[o > 0] = out[o > 0] / o[o > 0]                                             # :44-45
    return out
