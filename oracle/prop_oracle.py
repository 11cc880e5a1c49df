"""CPU oracle for flow-guided gradient propagation (TEST INFRASTRUCTURE — see oracle/fgt_oracle.py).

Dense numpy restatement of /root/reference/tool/get_flowNN_gradient.py:11-534 (Nonlocal=False, the
driver's setting) and of interp / BFconsistCheck / FBconsistCheck / consistCheck in
/root/reference/tool/utils/common_utils.py:149-254. The reference keeps per-hole-pixel lists
(`sub`, `flowNN[numPix,3,2]`); here the same state lives in dense [H,W,N] arrays, which is also the
layout the CUDA kernels use. cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) is restated by `remap_q32`
(OpenCV quantises sample coordinates to 1/32 pixel: sx = rint(x*32), tap = sx>>5, frac = (sx&31)/32;
third-party: opencv-python 4.13.0 in this image, pinned by tests against cv2.remap itself).
Pinned against the unmodified reference by tests/golden/prop_*.npz.
"""
import numpy as np

NO_NN = 99999.0


def remap_q32(img, x, y):
    """cv2.remap(img, x, y, INTER_LINEAR) for float32 img [H,W] or [H,W,C] and float32 coordinate
    arrays; zero outside the image; coordinates quantised to 1/32 px like OpenCV's fixed-point maps."""
    img = np.asarray(img, dtype=np.float32)
    H, W = img.shape[:2]
    sx = np.rint(x.astype(np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(y.astype(np.float32) * np.float32(32)).astype(np.int64)
    ix, iy = sx >> 5, sy >> 5
    fx = ((sx & 31).astype(np.float32) / np.float32(32))
    fy = ((sy & 31).astype(np.float32) / np.float32(32))

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[..., None] if img.ndim == 3 else ok, v, np.float32(0))

    w00 = (np.float32(1) - fy) * (np.float32(1) - fx)
    w01 = (np.float32(1) - fy) * fx
    w10 = fy * (np.float32(1) - fx)
    w11 = fy * fx
    if img.ndim == 3:
        w00, w01, w10, w11 = (w[..., None] for w in (w00, w01, w10, w11))
    return (tap(iy, ix) * w00 + tap(iy, ix + 1) * w01 + tap(iy + 1, ix) * w10 + tap(iy + 1, ix + 1) * w11).astype(
        np.float32)


def _pass(mask, flow_fwd, flow_bwd, direction, thres, nn, have, cuv, cmap):
    """One propagation pass. direction = -1: 'Forward Pass' of the reference (backward-flow neighbours,
    NN slot 0, get_flowNN_gradient.py:76-235); +1: 'Backward Pass' (slot 1, :241-370)."""
    H, W, N = mask.shape
    slot = 0 if direction < 0 else 1
    frames = range(1, N) if direction < 0 else range(N - 2, -1, -1)
    for t in frames:
        tn = t + direction
        ys, xs = np.nonzero(mask[:, :, t])
        if direction < 0:
            step, back = flow_bwd[:, :, :, t - 1], flow_fwd[:, :, :, t - 1]  # t->t-1, then t-1->t
        else:
            step, back = flow_fwd[:, :, :, t], flow_bwd[:, :, :, t]
        ny = ys.astype(np.float32) + step[ys, xs, 1]
        nx = xs.astype(np.float32) + step[ys, xs, 0]
        iy = np.round(ny).astype(np.int32)
        ix = np.round(nx).astype(np.int32)
        # round-trip consistency (BFconsistCheck / FBconsistCheck / consistCheck)
        ry = ny + remap_q32(back[:, :, 1], nx, ny)
        rx = nx + remap_q32(back[:, :, 0], nx, ny)
        diff = ((ry.astype(np.float64) - ys) ** 2 + (rx.astype(np.float64) - xs) ** 2) ** 0.5
        consist = diff < thres
        u_abs = np.abs(rx - xs.astype(np.float32)).astype(np.float64)
        v_abs = np.abs(ry - ys.astype(np.float32)).astype(np.float64)
        inb = (iy >= 0) & (iy < H - 1) & (ix >= 0) & (ix < W - 1)
        cy, cx = np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)
        known = inb & (mask[cy, cx, tn] == 0)
        # case 1: the flow neighbour is a known pixel
        c1 = known & consist
        nn[ys[c1], xs[c1], t, slot, 0] = ny[c1]
        nn[ys[c1], xs[c1], t, slot, 1] = nx[c1]
        nn[ys[c1], xs[c1], t, slot, 2] = tn
        have[ys[c1], xs[c1], t, slot] = 1
        cuv[ys[c1], xs[c1], slot, 0, t] = u_abs[c1]
        cuv[ys[c1], xs[c1], slot, 1, t] = v_abs[c1]
        # case 2: the neighbour is a hole pixel that already has a neighbour: chain with the rounding residue
        c2 = inb & ~known & (have[cy, cx, tn, slot] == 1) & consist
        ref_y = ny.astype(np.float64) - iy
        ref_x = nx.astype(np.float64) - ix
        ty = nn[cy, cx, tn, slot, 0] + ref_y
        tx = nn[cy, cx, tn, slot, 1] + ref_x
        tyi, txi = np.round(ty).astype(np.int64), np.round(tx).astype(np.int64)
        c2 &= (tyi >= 0) & (tyi < H - 1) & (txi >= 0) & (txi < W - 1)
        nn[ys[c2], xs[c2], t, slot, 0] = ty[c2]
        nn[ys[c2], xs[c2], t, slot, 1] = tx[c2]
        nn[ys[c2], xs[c2], t, slot, 2] = nn[cy[c2], cx[c2], tn, slot, 2]
        have[ys[c2], xs[c2], t, slot] = 1
        cuv[ys[c2], xs[c2], slot, 0, t] = np.maximum(u_abs[c2], np.abs(cuv[cy[c2], cx[c2], slot, 0, tn]))
        cuv[ys[c2], xs[c2], slot, 1, t] = np.maximum(v_abs[c2], np.abs(cuv[cy[c2], cx[c2], slot, 1, tn]))
        cmap[:, :, slot, t] = (cuv[:, :, slot, 0, t] ** 2 + cuv[:, :, slot, 1, t] ** 2) ** 0.5


def get_flownn_gradient(gradient_x, gradient_y, mask, flow_f, flow_b, consistency_thres=5.0, alpha=0.1):
    """Returns (gradient_x, gradient_y, mask_tofill) like the reference (inputs are not modified).
    gradient_*: [H,W,3,N] float32; mask: [H,W,N] bool; flow_*: [H,W,2,N-1] float32 (u, v)."""
    mask = mask.astype(bool)
    H, W, N = mask.shape
    nn = np.full((H, W, N, 2, 3), NO_NN, dtype=np.float64)
    have = np.full((H, W, N, 2), NO_NN, dtype=np.float64)
    have[mask, :] = 0
    cuv = np.zeros((H, W, 2, 2, N))
    cmap = np.zeros((H, W, 2, N))
    _pass(mask, flow_f, flow_b, -1, consistency_thres, nn, have, cuv, cmap)
    _pass(mask, flow_f, flow_b, +1, consistency_thres, nn, have, cuv, cmap)
    # ordered in-place interpolation (get_flowNN_gradient.py:378-435)
    cands = []
    for slot, order in ((0, range(N)), (1, range(N - 1, -1, -1))):
        gx, gy = gradient_x.copy(), gradient_y.copy()
        for s in order:
            ys, xs, ts = np.nonzero((nn[:, :, :, slot, 2] == s) & mask)
            if len(ys) == 0:
                continue
            px = nn[ys, xs, ts, slot, 1].astype(np.float32)
            py = nn[ys, xs, ts, slot, 0].astype(np.float32)
            gx[ys, xs, :, ts] = remap_q32(gx[:, :, :, s], px, py)
            gy[ys, xs, :, ts] = remap_q32(gy[:, :, :, s], px, py)
        cands.append((gx, gy))
    # confidence-weighted fusion (get_flowNN_gradient.py:440-532)
    out_x, out_y = gradient_x.copy(), gradient_y.copy()
    tofill = np.zeros((H, W, N), dtype=bool)
    for t in range(N):
        hv = np.stack([have[:, :, t, 0] == 1, have[:, :, t, 1] == 1], -1)
        nothave = ~hv & mask[:, :, t][..., None]
        anyv = hv[..., 0] | hv[..., 1]
        cm = np.exp(-cmap[:, :, :, t] / alpha)
        cm[nothave] = 0
        num = cm * hv
        den = num.sum(-1, keepdims=True)
        with np.errstate(invalid="ignore", divide="ignore"):
            wgt = np.where(den == 0, hv / np.maximum(hv.sum(-1, keepdims=True), 1), num / den)
        for c in range(3):
            fx = cands[0][0][:, :, c, t].astype(np.float64) * wgt[..., 0] + cands[1][0][:, :, c, t].astype(
                np.float64) * wgt[..., 1]
            fy = cands[0][1][:, :, c, t].astype(np.float64) * wgt[..., 0] + cands[1][1][:, :, c, t].astype(
                np.float64) * wgt[..., 1]
            out_x[anyv, c, t] = fx[anyv]
            out_y[anyv, c, t] = fy[anyv]
        tofill[:, :, t] = ~anyv & mask[:, :, t]
    return out_x, out_y, tofill
